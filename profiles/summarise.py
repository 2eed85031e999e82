#!/usr/bin/env python
"""Turn an ncu report into the text summary committed under profiles/.

    python profiles/summarise.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_full.md
    python profiles/summarise.py --launches gpurun_out/launches.csv > profiles/rNN_launches.md

Needs only the `ncu` CLI (no GPU): `--page raw` for the counters, `--page source` for the
per-instruction stall samples.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def summarise_report(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    print("# ncu --set full summary of `%s`\n" % rep.split("/")[-1])
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        print("## launch %d: `%s`\n" % (n, name))
        print("| metric | value | unit |\n|---|---|---|")
        traffic = 0.0
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("| %s | %s | %s |" % (k, r[i], units[i]))
                if k.startswith("dram__bytes_"):
                    traffic += to_bytes(r[i], units[i])
        print("| **DRAM traffic (read + write)** | %.1f | MB per launch |\n" % (traffic / 1e6))
    src = ncu_csv(rep, "source")
    # the source page repeats a 2-line header per kernel; take every data row with an address
    hdr = None
    stall_tot, total = defaultdict(int), 0
    insts = []
    for r in src:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        ix = {h: i for i, h in enumerate(hdr)}
        try:
            ns = int(r[ix["# Samples"]] or 0)
        except ValueError:
            continue
        total += ns
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stall_tot[h] += int(r[ix[h]] or 0)
                except ValueError:
                    pass
        insts.append((ns, r[ix["Instructions Executed"]], r[ix["Source"]]))
    if total:
        print("## warp stall samples (all launches in the report, %d samples)\n" % total)
        print("| reason | samples | share |\n|---|---|---|")
        for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]):
            if v:
                print("| %s | %d | %.1f%% |" % (k, v, 100.0 * v / total))
        print("\n## hottest SASS instructions\n")
        print("| samples | executed | instruction |\n|---|---|---|")
        for ns, ex, s in sorted(insts, key=lambda t: -t[0])[:25]:
            print("| %d | %s | `%s` |" % (ns, ex, s.strip()[:90]))


def summarise_launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)   # -> us
        name = r[ik].split("(")[0]
        if "native::" in name or "at_cuda_detail" in name or "at::" in name:
            name = "(torch kernels: synthetic-data setup, outside the timed region)"
        name = name[:110]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu launch list `%s` (gpu__time_duration.sum; cold-cache, serialised: compare SHARES)\n"
          % path.split("/")[-1])
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (name, n, t, t / n, 100.0 * t / tot))
    print("\ntotal %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--launches":
        summarise_launches(sys.argv[2])
    else:
        summarise_report(sys.argv[1])
