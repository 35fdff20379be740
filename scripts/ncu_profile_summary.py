"""Turn an ncu report into the committed summaries under profiles/:
   python scripts/ncu_profile_summary.py gpurun_out/prof.ncu-rep profiles/r1_<tag> [--traffic] [--issue]
writes <out>_raw_summary.json (+ profiles/traffic.json with --traffic, profiles/issue_roofline.json with --issue)
and <out>_source_summary.txt"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
launches = [{"launch": i, **{k: (r[idx[k]] + " " + units[idx[k]]).strip() for k in keys if k in idx}}
            for i, r in enumerate(data)]
json.dump({"what": "ncu --set full --clock-control none; bench workload C2 at full scale, one launch per camera "
                   "pose (0 start pose, 1 inside building, 2 open sky, 3 grazing terrain); ncu flushes caches "
                   "between replays, so DRAM figures are cold-cache", "report": rep, "launches": launches},
          open(out + "_raw_summary.json", "w"), indent=1)


def num(r, k):
    return float(r[idx[k]])


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


if "--traffic" in sys.argv:
    tot = [to_bytes(num(r, "dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]])
           + to_bytes(num(r, "dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]]) for r in data]
    json.dump({"dram_bytes_per_launch": sum(tot) / len(tot), "per_pose": tot, "source": out + "_raw_summary.json",
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of the render kernel, mean over the four poses"},
              open("profiles/traffic.json", "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:render", "--launch-skip",
                      "0", "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
txt = subprocess.run([sys.executable, "scripts/ncu_source_summary.py", "/tmp/_src.csv", "12"], stdout=subprocess.PIPE,
                     text=True).stdout
open(out + "_source_summary.txt", "w").write(txt)
print(json.dumps(launches[0], indent=1))

if "--issue" in sys.argv:
    # the instruction-issue roofline of the render kernel (what bounds it; bench.py copies it into its line):
    # issue slots used, lanes per issued instruction, warp instructions per traversal step.  A traversal step = one pop
    # (descent); its count is the execution count of the level-state store (STS) of the loop, from the source page.
    import re
    per = []
    for k in range(len(data)):
        srck = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:render",
                               "--launch-skip", str(k), "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
        rr = list(csv.reader(io.StringIO(srck)))
        h0 = next(i for i, r in enumerate(rr) if r and r[0] == "Address")
        hd = {c: i for i, c in enumerate(rr[h0])}
        tab = []
        for r in rr[h0 + 1:]:
            if r and r[0] == "Address":
                break
            if r and r[0].startswith("0x"):
                tab.append(r)
        ex = lambda r: int(float(r[hd["Instructions Executed"]] or 0))
        sts = max([ex(r) for r in tab if re.search(r"\bSTS\b", r[hd["Source"]])] or [0])
        tot = sum(ex(r) for r in tab)
        per.append({"warp_instructions": tot, "traversal_steps_warp": sts,
                    "warp_instructions_per_step": (tot / sts) if sts else None,
                    "issue_active_pct": num(data[k], "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "lanes_per_instruction": num(data[k], "smsp__thread_inst_executed_per_inst_executed.ratio")})
    ti = sum(p["warp_instructions"] for p in per)
    ts = sum(p["traversal_steps_warp"] for p in per)
    json.dump({"bound": "instruction issue", "unit": "fraction of issue slots",
               "issue_active": sum(p["issue_active_pct"] * p["warp_instructions"] for p in per) / ti / 100.0,
               "lanes_per_instruction": sum(p["lanes_per_instruction"] * p["warp_instructions"] for p in per) / ti,
               "lane_weighted_issue_utilisation": sum(p["issue_active_pct"] / 100.0 * p["lanes_per_instruction"] / 32.0
                                                      * p["warp_instructions"] for p in per) / ti,
               "warp_instructions_per_traversal_step": ti / ts if ts else None, "per_pose": per,
               "capture": out + "_raw_summary.json",
               "note": "the render kernel issues an instruction on `issue_active` of the cycles its sub-partitions are "
                       "active, with `lanes_per_instruction` of 32 lanes doing work; a traversal step (pop + descend + "
                       "expand) costs `warp_instructions_per_traversal_step` warp instructions all told (ray set-up, "
                       "shading and stores included)"}, open("profiles/issue_roofline.json", "w"), indent=1)
