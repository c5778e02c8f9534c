"""ncu counters -> JSON for bench.py's roofline (traffic, instruction counts), tied to the kernel source they were measured on.

  ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_extract.py raw.csv lde "<command that was profiled>"

Writes profiles/r02_ncu_<kind>.json: per kernel launch the duration, DRAM bytes and instruction counts, and for the LDE the
per-call totals (pass 1 + pass 2).  bench.py only uses the file while sha256(csrc/ntt.cuh) still matches."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6,
            "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)


def main():
    raw, kind = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
    launches = []
    for r in rows[2:]:
        e = {"kernel": r[idx["Kernel Name"]]}
        for w in want:
            if w in idx:
                v = num(r[idx[w]])
                if v is not None:
                    e[w] = v * unit_scale(units[idx[w]])
        launches.append(e)
    with open(os.path.join(ROOT, "ministark_b200", "csrc", "ntt.cuh"), "rb") as fh:
        sha = hashlib.sha256(fh.read()).hexdigest()
    out = {"source": f"ncu --set full --clock-control none, {os.path.basename(raw)} ({cmd})", "ntt_cuh_sha256": sha, "launches": launches}
    if kind == "lde":
        # one ms_coset_lde call = the launches of one pass-1 + pass-2 pair (the capture holds whole calls)
        ntt = [l for l in launches if "k_ntt" in l["kernel"]]
        calls = max(1, len(ntt) // 2)
        out["calls_captured"] = calls
        out["dram_bytes_per_call"] = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in ntt) / calls
        if all("smsp__thread_inst_executed.sum" in l for l in ntt):
            out["thread_inst_executed_per_call"] = sum(l["smsp__thread_inst_executed.sum"] for l in ntt) / calls
        else:
            out["thread_inst_executed_per_call"] = 32 * sum(l.get("smsp__inst_executed.sum", 0) for l in ntt) / calls
        out["warp_inst_executed_per_call"] = sum(l.get("smsp__inst_executed.sum", 0) for l in ntt) / calls
        out["kernel_time_s_per_call_under_ncu"] = sum(l.get("gpu__time_duration.sum", 0) for l in ntt) / calls
    path = os.path.join(ROOT, "profiles", f"r02_ncu_{kind}.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(path, {k: v for k, v in out.items() if k != "launches"})


if __name__ == "__main__":
    main()
