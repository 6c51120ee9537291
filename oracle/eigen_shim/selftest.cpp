// TEST INFRASTRUCTURE ONLY: C-ABI probes into the Eigen stand-in (oracle/eigen_shim/Eigen/Core) so that
// tests/test_eigen_shim.py can check its semantics against numpy / scipy, independently of the reference sources
// that are compiled against it (oracle/ref_build.py).
#include "Eigen/Dense"
#include "Eigen/Geometry"

extern "C" {

// C = A (r x k) * B (k x c), row-major in and out
void shim_matmul(const double* a, const double* b, int r, int k, int c, double* out) {
    Eigen::MatrixXd A(r, k), B(k, c);
    for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) A(i, j) = a[i * k + j];
    for (int i = 0; i < k; ++i) for (int j = 0; j < c; ++j) B(i, j) = b[i * c + j];
    Eigen::MatrixXd C = A * B;
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) out[i * c + j] = C(i, j);
}

// a script of view / initialiser operations on a 4x4 matrix; returns 16 row-major values + 8 extras
void shim_views(double* out24) {
    Eigen::Matrix4d m;
    m << 1, 2, 3, 4,
         5, 6, 7, 8,
         9, 10, 11, 12,
         13, 14, 15, 16;                                  // comma initialiser fills row by row
    Eigen::Matrix3d r3 = Eigen::Matrix3d::Identity() * 2.0;
    m.block(0, 0, 3, 3) = r3 + m.block(1, 1, 3, 3);       // block = expression of an overlapping block
    Eigen::Vector4d col = m.col(3);
    m.row(3) = col;                                      // a column assigned to a row view (transposed on assignment)
    Eigen::VectorXd v = Eigen::VectorXd::Zero(6);
    v.segment(1, 3) = m.col(0).segment(0, 3) * 0.5;
    v.tail(2) = Eigen::Vector2d(7, 8);
    Eigen::Matrix4d t = m.transpose();
    m(2, 1) += t(0, 3);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out24[4 * i + j] = m(i, j);
    for (int i = 0; i < 6; ++i) out24[16 + i] = v[i];
    Eigen::Vector4d a(1, 2, 3, 9), b(-2, 0.5, 4, 9);
    Eigen::Vector4d c = a.cross3(b);                      // cross product of the first three entries, w = 0
    out24[22] = c.dot(Eigen::Vector4d(1, 1, 1, 1));
    out24[23] = (a.cwiseMax(b) - a.cwiseMin(b)).squaredNorm() + a.segment(0, 3).norm() + m.data()[1];  // column-major data()
}

// quaternions are (w, x, y, z)
void shim_quat_mul(const double* a, const double* b, double* out) {
    Eigen::Quaterniond q = Eigen::Quaterniond(a[0], a[1], a[2], a[3]) * Eigen::Quaterniond(b[0], b[1], b[2], b[3]);
    out[0] = q.w(); out[1] = q.x(); out[2] = q.y(); out[3] = q.z();
}
void shim_quat_rotate(const double* q, const double* v, double* out) {
    Eigen::Vector3d r = Eigen::Quaterniond(q[0], q[1], q[2], q[3]) * Eigen::Vector3d(v[0], v[1], v[2]);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void shim_quat_slerp(const double* a, const double* b, double t, double* out) {
    Eigen::Quaterniond q = Eigen::Quaterniond(a[0], a[1], a[2], a[3]).slerp(t, Eigen::Quaterniond(b[0], b[1], b[2], b[3]));
    out[0] = q.w(); out[1] = q.x(); out[2] = q.y(); out[3] = q.z();
}
void shim_quat_misc(const double* a, double* out) {  // conjugate (4), inverse (4), normalized (4)
    Eigen::Quaterniond q(a[0], a[1], a[2], a[3]);
    Eigen::Quaterniond c = q.conjugate(), i = q.inverse(), n = q.normalized();
    const Eigen::Quaterniond qs[3] = {c, i, n};
    for (int k = 0; k < 3; ++k) { out[4 * k] = qs[k].w(); out[4 * k + 1] = qs[k].x(); out[4 * k + 2] = qs[k].y(); out[4 * k + 3] = qs[k].z(); }
}
void shim_from_two_vectors(const double* a, const double* b, double* out) {
    Eigen::Quaterniond q = Eigen::Quaterniond::FromTwoVectors(Eigen::Vector3d(a[0], a[1], a[2]), Eigen::Vector3d(b[0], b[1], b[2]));
    out[0] = q.w(); out[1] = q.x(); out[2] = q.y(); out[3] = q.z();
}
// x solving A x = b for a symmetric positive definite A (n x n, row-major)
void shim_solve(const double* a, const double* b, int n, double* out) {
    Eigen::MatrixXd A(n, n);
    Eigen::VectorXd B(n);
    for (int i = 0; i < n; ++i) { B[i] = b[i]; for (int j = 0; j < n; ++j) A(i, j) = a[i * n + j]; }
    Eigen::VectorXd x = A.ldlt().solve(B);
    for (int i = 0; i < n; ++i) out[i] = x[i];
}

}  // extern "C"
