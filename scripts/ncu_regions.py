"""Executed warp instructions of render_fast_kernel by REGION of the kernel source, from an ncu report taken with
--import-source on (SASS mapped to source lines by -lineinfo; inlined helpers count where their text is):
   python scripts/ncu_regions.py gpurun_out/prof.ncu-rep profiles/r2_ncu_<tag>_regions.json [launches]
Regions are found by marker comments / statements in the CURRENT source, so run it on the build the report came from.
Needs profiles/issue_roofline.json of the same capture (scripts/ncu_profile_summary.py --issue) for the step count."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out = sys.argv[1], sys.argv[2]
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 4


def load(k):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:render", "--launch-skip", str(k), "--launch-count", "1"], stdout=subprocess.PIPE,
                         text=True).stdout
    cur = hdr = None
    agg = {}
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            continue
        if r[0] == "Line No":
            hdr = {k: i for i, k in enumerate(r)}
            continue
        if hdr is None or cur is None or not r[0].strip().isdigit():
            continue
        try:
            w = int(float(r[hdr["Instructions Executed"]] or 0))
            t = int(float(r[hdr["Thread Instructions Executed"]] or 0))
        except (ValueError, IndexError):
            continue
        a = agg.setdefault((cur.split("/")[-1], int(r[0])), [0, 0])
        a[0] += w
        a[1] += t
    return agg


tot = collections.defaultdict(lambda: [0, 0])
for k in range(launches):
    for key, a in load(k).items():
        tot[key][0] += a[0]
        tot[key][1] += a[1]


def lines(f):
    return open(os.path.join(ROOT, "qubatron_b200", "csrc", f)).read().splitlines()


def find(L, s, start=0):
    for i in range(start, len(L)):
        if s in L[i]:
            return i + 1
    raise KeyError(s)


body, fast = lines("octree_trace_fast_body.inc"), lines("octree_trace_fast.cuh")
b_exp = find(body, "if (lbit == 1u)")
b_gen = find(body, "the general case, exactly as the reference orders it")
b_bt = find(body, "nothing here: continue from the deepest pending level")
f_slot0, f_slot1 = find(fast, "__device__ __forceinline__ int node_mask"), find(fast, "// per-ray division state")
f_sts0, f_sts1 = find(fast, "auto stack_store = [&]"), find(fast, "// CTA -> (view, shard tile, block inside the tile)")
f_loop, f_end = find(fast, "while (alive)"), find(fast, "a ray ended: consume it")
f_after = find(fast, "leaf lookup and shading of what the loop recorded")
f_begin = find(fast, "auto begin_ray = [&]() -> bool {")
f_begin_end = find(fast, "start = false;", f_begin)
f_compact = find(fast, "__device__ __forceinline__ bool base_cube_entry_compact")
f_compact_end = find(fast, "// Packed fp32 arithmetic (FADD2 / FMUL2) in the traversal body")
f_store = find(fast, "if (px < P.W && py < P.H)", f_after)
f_rcp0, f_rcp1 = find(fast, "__device__ __forceinline__ float rcp_refined"), find(fast, "// the operand range in which")
names = ["pixel set-up, store, tile cost", "shading behind the loop", "ray start (reciprocals, base-cube entry, root)",
         "ray end (record, next-ray decision)", "pop + descend", "loop control", "expand (common case)",
         "general ordering path", "backtrack"]
reg = collections.OrderedDict((k, [0, 0]) for k in names)
for (f, l), a in tot.items():
    if f == "octree_trace_fast_body.inc":
        r = "pop + descend" if l < b_exp - 8 else "expand (common case)" if l < b_gen - 1 else \
            "general ordering path" if l < b_bt else "backtrack"
    elif f == "octree_trace_fast.cuh":
        if f_compact <= l < f_compact_end or f_begin <= l < f_begin_end or f_rcp0 <= l < f_rcp1:
            r = "ray start (reciprocals, base-cube entry, root)"
        elif f_slot0 <= l < f_slot1 or f_sts0 <= l < f_sts1:
            r = "pop + descend"            # slot records, stack stores / loads (the loads belong to the backtrack: small)
        elif f_loop <= l < f_end:
            r = "loop control"
        elif f_end <= l < f_after - 1:
            r = "ray end (record, next-ray decision)"
        elif f_after - 1 <= l < f_store:
            r = "shading behind the loop"
        else:
            r = "pixel set-up, store, tile cost"
    else:
        r = "pixel set-up, store, tile cost"   # quat_rotate / normalize / unorm8 / shuffles / atomics of other headers
    reg[r][0] += a[0]
    reg[r][1] += a[1]
W = sum(a[0] for a in reg.values())
ir = json.load(open(os.path.join(ROOT, "profiles", "issue_roofline.json")))
steps = sum(p["traversal_steps_warp"] for p in ir["per_pose"][:launches])
res = {"what": "executed warp instructions by region of render_fast_kernel, %d bench poses (ncu source page of %s)" % (
    launches, os.path.basename(rep)), "warp_instructions": W, "traversal_steps_warp": steps, "regions": {}}
for k, a in reg.items():
    if a[0]:
        res["regions"][k] = {"share_pct": round(100.0 * a[0] / W, 2), "active_lanes": round(a[1] / a[0], 1),
                             "warp_instructions_per_step": round(a[0] / steps, 1)}
json.dump(res, open(out, "w"), indent=1)
for k, v in res["regions"].items():
    print("%-48s %5.1f %%  %4.1f lanes  %5.1f per step" % (k, v["share_pct"], v["active_lanes"], v["warp_instructions_per_step"]))
print("total per step %.1f" % (W / steps))
