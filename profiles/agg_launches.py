"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel family: launches, total ms, share."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = val * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    fam = name
    for pat, lab in ((r"tma_gemm_kernel<.*EpiSubTma", "tma_gemm_kernel<EpiSubTma> (trailing / inner update)"),
                     (r"tma_gemm_kernel<.*EpiGramTma", "tma_gemm_kernel<EpiGramTma> (Gram + recursion)"),
                     (r"tma_gemm_kernel<.*EpiGradTma", "tma_gemm_kernel<EpiGradTma> (gradient Gram pass)"),
                     (r"tma_gemm_kernel<.*EpiStoreTma", "tma_gemm_kernel<EpiStoreTma> (A^-1 = U U^T)"),
                     (r"tma_gemm_kernel<.*EpiScatterTma", "tma_gemm_kernel<EpiScatterTma> (panel solve + scatter)"),
                     (r"potf2_trtri", "potf2_trtri_kernel (diagonal block)"), (r"gemm_kernel<", "gemm_kernel cp.async (TRSM / small updates)")):
        if re.search(pat, name):
            fam = lab
            break
    tot[fam][0] += 1
    tot[fam][1] += ms
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>12s} {'share':>8s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {v[0]:8d} {v[1]:12.3f} {100 * v[1] / total:7.2f}%")
