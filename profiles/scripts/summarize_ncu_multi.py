"""Summarise a multi-kernel `ncu --set full` report: one line per captured launch with duration, DRAM bytes read/written
and achieved DRAM rate; optionally joined with the probe's JSON lines (algorithmic bytes per launch, in launch order).

    python profiles/scripts/summarize_ncu_multi.py gpurun_out/r02_hbm_kernels.ncu-rep [gpurun_out/r02_hbm_probe_plain.log] > profiles/r02_ncu_hbm_kernels.txt
"""
import csv, io, json, re, subprocess, sys

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9, "Tbyte": 1e12}
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] * 1e9


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    probe = []
    if len(sys.argv) > 2:
        for line in open(sys.argv[2]):
            if line.startswith("{"):
                probe.append(json.loads(line))
    # expand the probe records into launch order per kernel name
    expect = {}
    for rec in probe:
        expect.setdefault(rec["kernel"].split("<")[0], []).extend([rec] * rec["launches"])
    seen = {}
    print(f"# {rep}: ncu --set full --clock-control none; HBM peak (MEASURED_PEAKS.json) {PEAK/1e9:.1f} GB/s")
    print("# kernel | case | duration | algorithmic bytes -> GB/s (frac of measured peak) | DRAM read + write = traffic (x algorithmic) | dram throughput % | regs | grid x block")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"(\w+_kernel)", name)
        short = m.group(1) if m else name[:40]
        k = seen.get(short, 0); seen[short] = k + 1
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        pct = r[col["FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"]]
        regs = r[col["launch__registers_per_thread"]]
        grid, block = r[col["launch__grid_size"]], r[col["launch__block_size"]]
        rec = None
        for key, lst in expect.items():
            if key in short and k < len(lst):
                rec = lst[k]
        if rec:
            ab = rec["algorithmic_bytes"]
            print(f"{short:24s} | {rec['label']:42s} | {t*1e6:9.1f} us | {ab/1e6:10.2f} MB -> {ab/t/1e9:7.1f} GB/s ({ab/t/PEAK:5.3f}) | "
                  f"{rd/1e6:9.2f} + {wr/1e6:9.2f} = {(rd+wr)/1e6:10.2f} MB (x{(rd+wr)/ab:4.2f}) | {pct:>6s} % | {regs} | {grid} x {block}")
        else:
            print(f"{short:24s} | {'?':42s} | {t*1e6:9.1f} us | DRAM {rd/1e6:.2f} + {wr/1e6:.2f} MB -> {(rd+wr)/t/1e9:.1f} GB/s | {pct} % | {regs} | {grid} x {block}")


main()
