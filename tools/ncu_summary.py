"""Summarise an ncu report (`ncu --set full ... -o X`) as JSON: duration, DRAM bytes, L2 hit
rate, issue utilisation, registers, top stall reasons.   usage: ncu_summary.py X.ncu-rep [cell_updates]"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
  rep = sys.argv[1]
  cell_updates = float(sys.argv[2]) if len(sys.argv) > 2 else None
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                       text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  res = []
  for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    k = {"kernel": d["Kernel Name"], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for key in KEYS:
      if key in d:
        try:
          k[key] = {"value": float(d[key]), "unit": u[key]}
        except ValueError:
          pass
    stalls = []
    for h in hdr:
      if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        try:
          stalls.append((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "")
                         .replace("_per_issue_active.ratio", "")))
        except ValueError:
          pass
    k["stall_warps_per_issue_active"] = {n: round(v, 3) for v, n in sorted(stalls, reverse=True)[:8]}
    if cell_updates:
      scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
      rd = k.get("dram__bytes_read.sum"); wr = k.get("dram__bytes_write.sum")
      if rd and wr:
        b = rd["value"] * scale[rd["unit"]] + wr["value"] * scale[wr["unit"]]
        k["dram_bytes_per_launch"] = b
        k["dram_bytes_per_cell_update"] = b / cell_updates
      if "smsp__inst_executed.sum" in k:
        k["warp_instructions_per_cell_update"] = k["smsp__inst_executed.sum"]["value"] / cell_updates
    res.append(k)
  print(json.dumps(res, indent=1))


if __name__ == "__main__":
  main()
