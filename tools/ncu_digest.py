"""Digest an .ncu-rep (read offline with `ncu -i`): headline metrics, stall mix and the hottest SASS lines.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep [--json out.json] [--top 40]"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = {"kernel": d.get("Kernel Name", ("", ""))[0], "metrics": {k: d[k] for k in keys if k in d}}
st = []
for h, (v, u) in d.items():
    if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st.append((float(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1.0
out["stall_mix_pct"] = {h: round(100 * v / tot, 1) for v, h in sorted(st, reverse=True)[:8]}
for k, v in out["metrics"].items():
    print(f"{k:90s} {v[0]} {v[1]}")
print("stalls:", out["stall_mix_pct"])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h, data = rows[hi], rows[hi + 1:]
isrc, iss, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
iwx = h.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in h else None
tot = sum(int(r[iss] or 0) for r in data) or 1
hot = sorted(((int(r[iss] or 0), n, r) for n, r in enumerate(data)), reverse=True)[:top]
out["hot_sass"] = []
print(f"total samples {tot}; hottest SASS lines:")
for s_, n, r in sorted(hot, key=lambda t: t[1]):
    line = {"line": n, "sass": r[isrc][:80], "samples_pct": round(100 * s_ / tot, 2), "executed": r[iex], "smem_excess_wavefronts": (r[iwx] if iwx is not None else None)}
    out["hot_sass"].append(line)
    print(f"{n:5d} {r[isrc][:72]:72s} {100*s_/tot:6.2f}% exec={r[iex]} wfx={r[iwx] if iwx is not None else None}")
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
