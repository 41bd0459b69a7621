"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel time per kernel name and its share.

    python profiles/scripts/summarize_launches.py gpurun_out/r02_launches_c4s.csv > profiles/r02_launches_c4s_summary.txt
"""
import collections, csv, re, sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    by = collections.defaultdict(lambda: [0.0, 0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[4])[:110]
        t = float(r[14].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[13], 1e-6)
        by[name][0] += t; by[name][1] += 1
    tot = sum(v[0] for v in by.values())
    ours = sum(v[0] for k, v in by.items() if "at::" not in k and "cutlass" not in k and "nccl" not in k.lower())
    gemm = sum(v[0] for k, v in by.items() if "umma_gemm_kernel" in k)
    print(f"{len(rows)} launches, {tot:.1f} ms of kernel time; this repo's kernels {ours:.1f} ms, of which umma_gemm_kernel<*> {gemm:.1f} ms = {100 * gemm / ours:.1f} %")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{v[0]:9.2f} ms {100 * v[0] / tot:5.1f}% n={v[1]:5d} {k}")


main()
