#!/usr/bin/env python
"""Per-kernel-class summary of an `ncu --set full ... ; ncu -i X.ncu-rep --page raw --csv` dump -> profiles/r2_ncu_kernels.json (read by
bench.py for `roofline.traffic`) and a markdown table.  Usage: ncu_kernels_json.py raw.csv "source description" [out.json] [out.md]

Classes follow bench.py's KERNEL_BYTES keys: k_p2g launched with the fused G2P of the previous substep is "k_g2p2g" (same symbol as the
plain k_p2g: told apart by the executed instruction count), k_p2g_grad_g2p_grad is "k_p2g_grad+g2p_grad"."""
import csv
import json
import sys
import collections

csv.field_size_limit(10 ** 9)
raw, source = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_ncu_kernels.json"
out_md = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
for h, i in list(ix.items()):           # "FBSP.TriageCompute.dram__throughput..." is also reachable by its bare metric name
    if "." in h and h.split(".", 2)[-1] not in ix and h[0].isupper():
        ix.setdefault(h.split(".", 2)[-1], i)


def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError, IndexError):
        return None


def unit_scale(k):
    u = rows[1][ix[k]] if k in ix else ""
    return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u, 1.0)


p2g_inst = [num(r, "smsp__inst_executed.sum") for r in data if r[ix["Kernel Name"]].lstrip("void ").startswith("k_p2g<")]
p2g_min = min(p2g_inst) if p2g_inst else 0


def klass(r):
    n = r[ix["Kernel Name"]]
    n = n[5:] if n.startswith("void ") else n
    if n.startswith("k_p2g_grad_g2p_grad"): return "k_p2g_grad+g2p_grad"
    if n.startswith("k_p2g_grad"): return "k_p2g_grad"
    if n.startswith("k_p2g<"): return "k_g2p2g" if num(r, "smsp__inst_executed.sum") > 1.15 * p2g_min else "k_p2g"
    for k in ("k_g2p_grad", "k_g2p", "k_grid_op", "k_grid_grad", "k_contact_grad_sparse", "k_contact_grad", "k_contact", "k_ckpt_copy"):
        if n.startswith(k): return k
    return n.split("(")[0].split("<")[0]


M = {  # output key -> (ncu metric, scale by unit?)
    "time_us": ("gpu__time_duration.sum", True), "dram_read": ("dram__bytes_read.sum", True), "dram_write": ("dram__bytes_write.sum", True),
    "dram_pct_of_peak": ("dram__throughput.avg.pct_of_peak_sustained_elapsed", False),
    "lts_red_sectors": ("lts__t_sectors_op_red.sum", False), "lts_red_sectors_from_sm": ("lts__t_sectors_srcunit_tex_op_red.sum", False),
    "lts_throughput_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
    "l1_hit_pct": ("l1tex__t_sector_hit_rate.pct", False), "l1_data_pipe_pct": ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", False),
    "smem_wavefronts": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
    "warp_inst": ("smsp__inst_executed.sum", False), "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", False), "fma_pipe_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
    "registers": ("launch__registers_per_thread", False), "grid": ("launch__grid_size", False), "block": ("launch__block_size", False),
    "cycles": ("sm__cycles_elapsed.max", False),
}
acc = collections.defaultdict(lambda: collections.defaultdict(list))
for r in data:
    c = klass(r)
    for k, (m, sc) in M.items():
        v = num(r, m)
        if v is not None:
            acc[c][k].append(v * (unit_scale(m) if sc else 1.0))
kern = {}
for c, d in acc.items():
    e = {k: sum(v) / len(v) for k, v in d.items()}
    e["launches_captured"] = len(d["time_us"])
    e["time_us"] = e["time_us"] * 1e6
    e["dram_bytes"] = e.get("dram_read", 0.0) + e.get("dram_write", 0.0)
    # L2 reduction (atomic) throughput: 32-byte sectors of RED traffic per second while the kernel runs
    if e.get("lts_red_sectors"):
        e["lts_red_GBps"] = e["lts_red_sectors"] * 32 / (e["time_us"] * 1e-6) / 1e9
    e["dram_GBps"] = e["dram_bytes"] / (e["time_us"] * 1e-6) / 1e9
    e["dram_pct_of_peak"] = 100.0 * e["dram_GBps"] / 6467.7      # against the measured copy bandwidth (MEASURED_PEAKS.json)
    kern[c] = e
json.dump({"source": source, "note": "per launch, averaged over the captured launches of each class; ncu replays kernels cold and serialised: shares, not absolute times, compare with the bench",
           "kernels": kern}, open(out_json, "w"), indent=1)
lines = ["| kernel | launches | us | DRAM MB (R+W) | DRAM % of 6467.7 GB/s | L2 RED sectors | RED GB/s | L1 hit % | L1 data pipe % | warp inst (M) | issue % | warps active % | regs |", "|" + "---|" * 13]
for c, e in sorted(kern.items(), key=lambda kv: -kv[1]["time_us"] * kv[1]["launches_captured"]):
    lines.append("| `%s` | %d | %.1f | %.1f (%.1f + %.1f) | %.1f | %.3g | %.0f | %.1f | %.1f | %.2f | %.1f | %.1f | %d |" % (
        c, e["launches_captured"], e["time_us"], e["dram_bytes"] / 1e6, e.get("dram_read", 0) / 1e6, e.get("dram_write", 0) / 1e6, e.get("dram_pct_of_peak", 0),
        e.get("lts_red_sectors", 0), e.get("lts_red_GBps", 0), e.get("l1_hit_pct", 0), e.get("l1_data_pipe_pct", 0), e.get("warp_inst", 0) / 1e6,
        e.get("issue_active_pct", 0), e.get("warps_active_pct", 0), int(e.get("registers", 0))))
md = "\n".join(lines)
print(md)
if out_md:
    open(out_md, "w").write("ncu --set full, %s\n\n%s\n" % (source, md))
