"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into profiles/: key counters per capture + a JSON that
bench.py reads for roofline.traffic.  Usage: python profiles/scripts/summarize_ncu.py r01 gpurun_out/r01_*_full.ncu-rep"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        short = h.split("TriageCompute.")[-1]
        if short in KEYS or h in ("Kernel Name",):
            try:
                d[short] = (float(v.replace(",", "")), u)
            except ValueError:
                d[short] = (v, u)
    return d


def main():
    tag, paths = sys.argv[1], sys.argv[2:]
    lines, js = [], {}
    for p in paths:
        d = load(p)
        name = p.split("/")[-1].replace(".ncu-rep", "")
        t = d["gpu__time_duration.sum"][0] * UNIT[d["gpu__time_duration.sum"][1]]
        rd = d["dram__bytes_read.sum"][0] * UNIT[d["dram__bytes_read.sum"][1]]
        wr = d["dram__bytes_write.sum"][0] * UNIT[d["dram__bytes_write.sum"][1]]
        lines.append(f"== {name}: duration {t*1e3:.3f} ms | DRAM read {rd/1e9:.3f} GB + write {wr/1e9:.3f} GB = {(rd+wr)/1e9:.3f} GB "
                     f"-> {(rd+wr)/t/1e12:.2f} TB/s")
        for k in KEYS:
            if k in d:
                lines.append(f"   {k:80s} {d[k][0]} {d[k][1]}")
        js[name] = {"seconds": t, "dram_bytes_read": rd, "dram_bytes_write": wr}
    open(f"profiles/{tag}_ncu_full_summary.txt", "w").write("\n".join(lines) + "\n")
    json.dump(js, open(f"profiles/{tag}_ncu_traffic.json", "w"), indent=1)
    print("\n".join(l for l in lines if l.startswith("==")))


main()
