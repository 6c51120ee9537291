"""Per-kernel counts of the Blackwell-native SASS instructions in libsimstep.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel"):  cuobjdump -sass amp_extensions_b200/csrc/libsimstep.so | python tools/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import sys,re,collections,subprocess
cur=None; counts=collections.OrderedDict()
pat=re.compile(r'\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|UTCBAR|LDTM|STTM|HMMA|FFMA2|FMUL2|FADD2)\b')
for line in sys.stdin:
    m=re.search(r'Function : (\S+)', line)
    if m:
        cur=m.group(1); counts[cur]=collections.Counter(); continue
    if cur:
        for k in pat.findall(line): counts[cur][k]+=1
names=list(counts)
dem=subprocess.run(['c++filt']+names,capture_output=True,text=True).stdout.strip().split("\n")
print("# cuobjdump -sass amp_extensions_b200/csrc/libsimstep.so: Blackwell-native instructions per kernel")
print("# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,")
print("# FFMA2 / FMUL2 / FADD2 = packed fp32; HMMA (legacy mma.sync) must not appear")
for k,d in zip(names,dem):
    v=counts[k]
    if not v: continue
    name=re.sub(r'\(.*','',d).replace('void simstep::','')
    print(f"{name[:80]:80s} "+" ".join(f"{a}={b}" for a,b in sorted(v.items())))
