"""Read one `ncu --set full` capture of decode_persist_kernel and write profiles/r01_traffic.json (DRAM bytes per launch)
plus a compact CSV summary of the metrics DESIGN.md quotes.  Usage: ncu_traffic.py <report.ncu-rep> <batch> <mode> <out_prefix>"""
import csv, io, json, subprocess, sys
rep, batch, mode, out = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
def get(name):
    i = col.get(name)
    if i is None:
        return None, None
    return vals[i].replace(",", ""), units[i]
def to_bytes(v, u):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
rd, rdu = get("dram__bytes_read.sum"); wr, wru = get("dram__bytes_write.sum")
dur, duru = get("gpu__time_duration.sum")
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_op_shared_ld.sum", "l1tex__t_bytes.sum"]
with open(out + ".csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for k in keep:
        if k in col:
            w.writerow([k, units[col[k]], vals[col[k]]])
    for h in hdr:
        if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith(".ratio"):
            w.writerow([h, units[col[h]], vals[col[h]]])
total = to_bytes(rd, rdu) + to_bytes(wr, wru)
json.dump({"kernel": "decode_persist_kernel", "batch": batch, "mode": mode, "dram_bytes_per_launch": total,
           "dram_read_bytes": to_bytes(rd, rdu), "dram_write_bytes": to_bytes(wr, wru), "duration": dur + " " + duru,
           "source": "ncu --set full --clock-control none, one launch (cold-cache, serialised); " + rep.split("/")[-1]},
          open(out + ".json", "w"), indent=1)
print(open(out + ".json").read())
