"""CPU oracle of the hot path — TEST INFRASTRUCTURE ONLY.

oracle/ holds restatements of the reference's algorithms used to check the CUDA path:

  milo_oracle.py       torch-CPU fp32 / numpy float64 restatement of milo/milo/{dynamics,datasets,
                       linear_cost}.py and gym-simenv's SimEnv.step/is_done.  Pinned bit-exactly against
                       golden vectors produced by the reference's own code (tests/golden/make_golden.py).
  imitation_oracle.py  numpy float64 restatement of DeepMimicCore's CalcRewardImitate and the kinematic
                       pipeline under it.  The reference C++ needs Eigen 3.3.7 + Bullet 2.88 + SWIG and
                       cannot be built here: PARITY UNPINNED, anchored on closed-form known answers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  The product (amp_extensions_b200) never does.
"""
