"""SASS evidence of the shipped library (no GPU needed): instruction counts per kernel for the mnemonics that prove
tcgen05 / TMEM / TMA / mbarrier / cluster / DMMA use.  python scripts/sass_evidence.py > profiles/<round>_sass_evidence.txt"""
import collections
import os
import re
import subprocess

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyvbmc_b200", "csrc", "libvbmc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "UCGABAR", "MEMBAR.SYS", "DMMA", "FFMA2", "FFMA", "DFMA", "MUFU.EX2",
        "LDGSTS", "ACQBULK", "UTCATOMSWS", "ERRBAR", "SHFL", "BAR.SYNC", "LDS", "STS", "LDG", "STG")
KERNELS = ("entmc_kernel_tcILi20ELb1", "entmc_tc_gen_kernelILi20ELb1", "entmc_kernel_smallILi20ELb1ELb1ELb1", "entmc_kernel_wILi20ELb1ELb1ELb1",
           "gplj_kernel_wILi20ELb1", "tail_kernel", "gppred_kernelILi20", "vmul_kernel", "gram4_kernel", "adam_update_prepare_kernel",
           "theta_prepare_kernel", "sieve_kernel")
print("# SASS evidence (cuobjdump -sass libvbmc_b200.so, sm_100a), instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA 1-D),")
print("# SYNCS = mbarrier, UCGABAR = cluster barrier, DMMA = mma.sync.m8n8k4.f64, LDGSTS = cp.async, MEMBAR.SYS = system-scope fence")
cur, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + ".") or op.startswith(w + "_"):
                counts[cur][w if w in ("FFMA", "DFMA", "LDS", "STS", "LDG", "STG", "SHFL", "DMMA", "FFMA2", "LDGSTS") else op] += 1
                break
for kname in KERNELS:
    for fn in counts:
        if kname in fn:
            print(f"\n## {fn}")
            for op, n in counts[fn].most_common():
                print(f"{n:7d} {op}")
