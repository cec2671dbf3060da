"""Summarise an `ncu --set full` capture of ONE bf16 forward (25 kernels) into the files kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv    > raw.csv
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > src.csv   (optional, for the stall table)
    python tools/ncu_summarize.py raw.csv profiles/rNN_ncu_conv_tc_summary.csv [src.csv profiles/rNN_ncu_top_stalls.txt]

The summary's header records the source hash of the library that is loaded when this script runs (wunet_version(): run it
against the same build the capture was taken from); bench.py quotes `roofline.traffic` from the summary only when that hash
equals the running library's.

The level names follow the launch order of wunet_forward (enc0..enc11, middle, dec0..dec11); the per-launch times are
cold-cache and serialised, so only the SHARES are comparable with the bench's event timings.
"""
import csv
import sys


def col(header, name):
    return header.index(name)


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    h, data = rows[0], rows[2:]  # rows[1] is the units line
    n = (len(data) - 1) // 2
    names = [f"enc{i}" for i in range(n)] + ["middle"] + [f"dec{i}" for i in range(n)]
    keys = [
        ("dur_us", "gpu__time_duration.sum"),
        ("dram_rd_MB", "dram__bytes_read.sum"),
        ("dram_wr_MB", "dram__bytes_write.sum"),
        ("tensor_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"),
        ("grid", "launch__grid_size"),
        ("block", "launch__block_size"),
        ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ]
    units = rows[1]
    idx = [(k, col(h, m)) for k, m in keys]
    total = 0.0
    build = "unknown"
    try:
        import os
        import re
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from wave_u_net_for_speech_enhancement_b200 import _lib
        m = re.search(r"src ([0-9a-f]+)", _lib.load().wunet_version().decode())
        build = m.group(1) if m else "unknown"
    except Exception:  # noqa: BLE001
        pass
    with open(out, "w") as f:
        f.write("# build %s\n" % build)
        f.write("# ncu --set full --clock-control none of the %d kernels of ONE bf16 forward (B=256, T=16384)\n" % len(data))
        f.write("# per-launch times are cold-cache and serialised: use SHARES. dram_* in MB; tensor_pct = "
                "sm__pipe_tensor_subpipe_hmma_cycles_active % of peak\n")
        f.write("level," + ",".join(k for k, _ in keys) + "\n")
        for name, r in zip(names, data):
            vals = []
            for k, i in idx:
                v = float(r[i].replace(",", ""))
                u = units[i]
                if k.startswith("dram_") and k.endswith("MB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
                if k == "dur_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[u]
                    total += v
                if k == "smem_dyn_KB":
                    v *= {"byte": 1e-3, "Kbyte": 1.0}[u.split("/")[0]]
                vals.append("%g" % v if k in ("regs", "grid", "block") else "%.6f" % v)
            f.write(name + "," + ",".join(vals) + "\n")
        f.write("# sum of durations: %.1f us\n" % total)
    if len(sys.argv) >= 5:
        stalls(sys.argv[3], sys.argv[4], names)


def stalls(src, out, names, top=12):
    rows = list(csv.reader(open(src)))
    # the source page is a sequence of per-kernel tables: a "Kernel Name" line, a header line, then one row per SASS
    # instruction; ncu emits every table twice (two views of the same SASS), so keep every other one
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"h": None, "rows": []}
            tables.append(cur)
        elif cur is not None and r:
            if cur["h"] is None:
                cur["h"] = r
            else:
                cur["rows"].append(r)
    if len(tables) == 2 * len(names):
        tables = tables[::2]
    with open(out, "w") as f:
        f.write("# ncu --page source: top stall sites by warp-sample count. Source mapping via -lineinfo; listed as SASS.\n")
        for name, t in zip(names, tables):
            h = t["h"]
            try:
                isrc = h.index("Source")
                isam = h.index("# Samples")
                iexe = [i for i, c in enumerate(h) if c.strip() == "Instructions Executed"][0]
            except (ValueError, IndexError):
                continue
            stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            tot = sum(float(r[isam] or 0) for r in t["rows"])
            if tot < 3000:
                continue
            mix = sorted(((sum(float(r[i] or 0) for r in t["rows"]) / tot, c) for i, c in stall_cols), reverse=True)[:7]
            f.write("\n## %s: %d samples, %d SASS instructions\n" % (name, tot, len(t["rows"])))
            f.write("stall mix: " + ", ".join("%s=%.1f%%" % (c, 100 * v) for v, c in mix) + "\n")
            for r in sorted(t["rows"], key=lambda r: -float(r[isam] or 0))[:top]:
                dom = max(stall_cols, key=lambda ic: float(r[ic[0]] or 0))[1]
                f.write("  %5.1f%%  executed=%9s  %-70s %s\n" % (100 * float(r[isam] or 0) / tot, r[iexe], r[isrc][:70], dom))


if __name__ == "__main__":
    main()
