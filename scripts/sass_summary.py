"""Per-kernel SASS instruction counts of libmcpilco_b200.so (cuobjdump -sass): the mnemonics that prove the Blackwell-native paths
(UTCIMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA, UTCBAR = tcgen05.commit, DMMA = FP64 tensor pipe, SYNCS = mbarrier, UCGABAR =
cluster barrier) plus a short excerpt around the first tensor instruction.  usage: python scripts/sass_summary.py > profiles/rNN_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mc-pilco_b200", "mcpilco_b200", "libmcpilco_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1); kernels[cur] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and cur:
        kernels[cur].append(m.group(1).strip())
def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTCIMMA", "UTCBAR", "LDTM", "UTMALDG", "UTCATOMSWS", "UCGABAR", "SYNCS", "DMMA", "DFMA", "LDGSTS", "I2F.F64", "MUFU", "BAR.SYNC", "ELECT"]
want = sys.argv[1:] or ["ozaki_mma_kernel", "dgemm_tma_kernel", "small_gemm_kernel", "posterior_reduce_fast_kernelILi6ELi2", "cov_slice_kernelILi6ELi2",
                        "ozaki_slice_kernel", "posterior_tri_reduce_kernelILi6", "small_step_kernelILi6ELi2ELb1", "rollout_bwd_kernelILi8ELi2"]
print("# cuobjdump -sass mc-pilco_b200/mcpilco_b200/libmcpilco_b200.so (sm_100a): instruction counts per kernel")
for name, ins in kernels.items():
    if not any(w in name for w in want):
        continue
    cnt = collections.Counter()
    for i in ins:
        op = i.split()[1] if i.startswith("@") and len(i.split()) > 1 else i.split()[0]
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
    print("\n## %s\n   %d instructions; %s" % (demangle(name)[:160], len(ins), ", ".join("%s %d" % (k, cnt[k]) for k in KEYS if cnt[k])))
    idx = next((j for j, i in enumerate(ins) if re.search(r"\b(UTCIMMA|DMMA)", i)), None)
    if idx is not None:
        for i in ins[max(0, idx - 6):idx + 6]:
            print("      " + i[:150])
