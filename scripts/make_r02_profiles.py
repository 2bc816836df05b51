"""Builds the round-2 files under profiles/ from the ncu reports and bench lines a GPU run left in gpurun_out/.

  python scripts/make_r02_profiles.py <final bundle dir> <single-search dir v2> <single-search dir v1> <cfg3 dir> <objs dir> [final objs subdir]

Every number in profiles/r02_*.md comes out of an .ncu-rep through `ncu -i ... --page raw/source --csv`
(scripts/ncu_lines.py joins the source page with `nvdisasm -g` line info of the object that was profiled)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
final, single2, single1, cfg3, objs = sys.argv[1:6]
final_objs = sys.argv[6] if len(sys.argv) > 6 else "final"  # objects (and sources) of the build the final bundle ran
P = os.path.join(ROOT, "profiles")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue")]


def table(reps):
    """reps: [(label, path)] -> markdown table metric x report"""
    cols = []
    for label, path in reps:
        hdr, units, rows = raw(path)
        r = rows[0]
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k, name in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                d[name] = f"{v} {u}".strip()
        cols.append((label, d))
    names = ["kernel"] + [n for _, n in KEYS]
    out = "| metric | " + " | ".join(l for l, _ in cols) + " |\n|---|" + "---|" * len(cols) + "\n"
    for n in names:
        out += f"| {n} | " + " | ".join(d.get(n, "-") for _, d in cols) + " |\n"
    return out


def lines(rep, obj, kern, mangled, top=16):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, obj, kern, str(top), mangled],
                         capture_output=True, text=True).stdout
    return "```\n" + out.rstrip() + "\n```\n"


def src_line(path, n):
    try:
        return open(path).read().split("\n")[n - 1].strip()
    except Exception:
        return ""


# ---- single search ----
with open(os.path.join(P, "r02_ncu_astar_single.md"), "w") as f:
    f.write("# ncu --set full of ONE A* search (256-room maze, 86 725 explored nodes; open field, 53 411)\n\n"
            "`ncu --set full --import-source on --clock-control none -k regex:k_astar --launch-skip 1 --launch-count 1 "
            "python scripts/profile_single.py <maze|field> <warp|thread> 1`.  Under ncu the L2 is flushed before every "
            "replay pass, so every edge record of the search misses L2 once (16.8 MB of DRAM reads = the record table); "
            "the un-profiled run of the same search is ~15 % faster (scripts/astar_latency.py).\n\n")
    f.write(table([("warp kernel, first version (loads into registers)", os.path.join(single1, "ncu_warp_maze1.ncu-rep")),
                   ("warp kernel, cp.async staging (maze)", os.path.join(single2, "ncu_warp_maze1.ncu-rep")),
                   ("thread kernel `<1,1536,2>` (maze)", os.path.join(single1, "ncu_thread_maze1.ncu-rep")),
                   ("warp kernel, cp.async staging (field, 1472 entries on chip)", os.path.join(single2, "ncu_warp_field1.ncu-rep"))]))
    f.write("\nPer explored node: 397 -> 340 warp instructions (warp kernel), 411 (thread kernel).\n\n"
            "## The false scoreboard wait in front of the pop (first version)\n\n"
            "25.5 % of all stall samples of the first version sat on ONE instruction, `IMAD.MOV.U32 R31, RZ, RZ, R36` — the "
            "first use of the `LDS.128` that reads the heap's last entry at the start of `pop()` — with reason "
            "`long_scoreboard`: ptxas gave that LDS the scoreboard of the three outstanding `LDG`s of the records / distance "
            "row, so the pop waited the whole L2 round trip it was meant to overlap.  With `cp.async` (no destination "
            "register, no scoreboard) the wait moves to `cp.async.wait_group` BEHIND the pop and shrinks to the part of "
            "the round trip the (short, maze) pop does not cover:\n\n")
    f.write("Warp kernel with cp.async staging, maze, top source lines by stall samples "
            "(`k_astar_warp.cu:183` at that commit = `cp.async.wait_group 0`):\n\n")
    f.write(lines(os.path.join(single2, "ncu_warp_maze1.ncu-rep"), os.path.join(objs, "r2c", "k_astar_warp.o"), "k_astar_warp",
                  "k_astar_warpILj1472ELj2ELi3ELb1E"))
    f.write("\nSame kernel on the open field with 1 472 on-chip entries (`:109` / `:47` = the spilled levels of the pop: "
            "loads of heap entries from global memory and the comparison that waits for them; the 13 824-entry shape "
            "added afterwards keeps a lone search's whole open list on chip: 2.06 -> 1.61 us per node):\n\n")
    f.write(lines(os.path.join(single2, "ncu_warp_field1.ncu-rep"), os.path.join(objs, "r2c", "k_astar_warp.o"), "k_astar_warp",
                  "k_astar_warpILj1472ELj2ELi0ELb1E"))
    f.write("\nThread kernel, maze (`k_astar.cu:295` = the pin behind the pop: the same L2 round trip of the records):\n\n")
    f.write(lines(os.path.join(single1, "ncu_thread_maze1.ncu-rep"), os.path.join(objs, "r2d", "k_astar.o"), "k_astar<",
                  "k_astarILj1ELj1536ELj2ELj2ELi3ELb1E"))
    pair = os.path.join(final, "ncu_pair_maze1.ncu-rep")
    if os.path.exists(pair):
        f.write("\n## Two warps per query (search warp + heap warp), the final kernel\n\n"
                "`... python scripts/profile_single.py maze pair 1`: both warps of the pair are in the one capture (64 threads), so "
                "'warp instructions' counts both and every percentage below is of the two warps together.  `barrier` stalls are "
                "the hand-overs: the search warp waiting for the pushes (the line after the second `bar.sync` of `peek`), the heap "
                "warp waiting for the batch (`heap_warp_loop`, first read of the mailbox).\n\n")
        f.write(table([("warp pair (maze)", pair)]))
        hdr, units, rows = raw(pair)
        inst = float(rows[0][hdr.index("smsp__inst_executed.sum")])
        f.write(f"\nPer explored node: {inst / 86725:.0f} warp instructions over the two warps.  Top source lines by stall samples:\n\n")
        f.write(lines(pair, os.path.join(objs, final_objs, "k_astar_warp.o"), "k_astar_warp",
                      "k_astar_warpILj13760ELj1ELj1ELi3ELb1ELb1E", 20))

# ---- loaded ----
rep = os.path.join(final, "ncu_astar_loaded.ncu-rep")
if os.path.exists(rep):
    hdr, units, rows = raw(rep)
    with open(os.path.join(P, "r02_ncu_astar_loaded_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(hdr)
        w.writerow(units)
        for r in rows:
            w.writerow(r)
    log = open(os.path.join(final, "ncu_loaded.log")).read() if os.path.exists(os.path.join(final, "ncu_loaded.log")) else ""
    with open(os.path.join(P, "r02_ncu_astar_loaded.md"), "w") as f:
        f.write("# ncu --set full of a fully loaded `k_astar<16,40,5,quad,forest,SIMPLE>` launch\n\n"
                "`ncu --set full --import-source on --clock-control none -k regex:k_astar --launch-skip 1 --launch-count 1 "
                "python scripts/profile_loaded.py 128 250000 47360`: 128x128-room maze (32 767 polygons), 250 000 queries on "
                "47 360 arena slots (the bench's slot count; 0.29 MB per slot, so kernel replay can save / restore the "
                "arena).  3.978 G explored nodes per launch.\n\n```\n" + log.strip()[-600:] + "\n```\n\n")
        f.write(table([("loaded launch", rep)]))
        r = rows[0]
        inst = float(r[hdr.index("smsp__inst_executed.sum")])
        f.write(f"\nPer explored node: {inst / 3977993509:.1f} warp instructions; DRAM "
                f"{(float(r[hdr.index('dram__bytes_read.sum')]) + float(r[hdr.index('dram__bytes_write.sum')])) * 1e9 / 3977993509:.1f} B "
                "(the units of the two DRAM rows above are Gbyte).\n\nTop source lines by stall samples:\n\n")
        f.write(lines(rep, os.path.join(objs, "final", "k_astar.o"), "k_astar<", "k_astarILj16ELj40ELj5ELj2ELi3ELb1E", 24))

# ---- config 3: avoidance / follow ----
a, fo = os.path.join(cfg3, "ncu_avoid_cfg3.ncu-rep"), os.path.join(cfg3, "ncu_follow_cfg3.ncu-rep")
if os.path.exists(a):
    with open(os.path.join(P, "r02_ncu_avoid_follow_cfg3.md"), "w") as f:
        f.write("# ncu --set full of `k_avoidance` (one 131 072-agent chunk) and `k_follow` (1 M agents) at config 3\n\n"
                "`ncu --set full -k regex:k_avoidance|k_follow --launch-skip 4 --launch-count 1 python bench.py --config 3 "
                "--steps 3 --warmup 2` (open field, 1 M agents, steady frame; scratch interleaved by agent, portal table, agents "
                "taken in cell order / longest corridor first).\n\n")
        f.write(table([("k_avoidance", a), ("k_follow", fo)]))
        f.write("\nk_avoidance, top source lines by stall samples (`lmb_device.cuh:29-32` = the first use of a gathered "
                "neighbour record, the `k_avoid.cu` lines around 1060 = `select_next`, the scan of the candidate list in the agent's "
                "interleaved scratch that yields the neighbours in (distance^2, index) order, `k_avoid.cu:53` = a `Strided` "
                "scratch access):\n\n")
        f.write(lines(a, os.path.join(objs, final_objs, "k_avoid.o"), "k_avoidance", "k_avoidance", 12))

# ---- plain copies ----
for src, dst in (("ncu_launches_bench.csv", "r02_ncu_launches_bench.csv"), ("ncu_dram_bench.csv", "r02_ncu_dram_bench.csv"),
                 ("bench_cfg2.json", "r02_bench_1gpu.json"), ("bench_ref.json", "r02_bench_reference_arm.json")):
    if os.path.exists(os.path.join(final, src)):
        shutil.copy(os.path.join(final, src), os.path.join(P, dst))
# every bench line of the other configs, one per line
names = ["bench_cfg3.json", "bench_cfg4_cell1_100k.json", "bench_cfg4_100k.json", "bench_cfg4_500k.json",
         "bench_cfg5_1000.json", "bench_cfg5_10000.json", "bench_cfg5_50000.json", "config1.json"]
with open(os.path.join(P, "r02_configs.jsonl"), "w") as f:
    for n in names:
        path = os.path.join(final, n)
        if os.path.exists(path) and os.path.getsize(path) > 2:
            try:
                d = json.loads(open(path).read().strip().split("\n")[-1])
            except Exception:
                continue
            d["_file"] = n
            f.write(json.dumps(d) + "\n")
if os.path.exists(os.path.join(final, "latency.jsonl")):
    shutil.copy(os.path.join(final, "latency.jsonl"), os.path.join(P, "r02_astar_latency.jsonl"))
    # the 1 M-polygon tiled world takes five GPU-minutes: its rows may come from an earlier run (R02_TILED_LATENCY)
    extra = os.environ.get("R02_TILED_LATENCY")
    have_tiled = any('"tiled"' in l for l in open(os.path.join(P, "r02_astar_latency.jsonl")))
    if extra and os.path.exists(extra) and not have_tiled:
        with open(os.path.join(P, "r02_astar_latency.jsonl"), "a") as f:
            for l in open(extra):
                if '"mesh": "tiled"' in l:
                    d = json.loads(l)
                    d["note"] = "measured one commit before the final pair kernel (prefetch by the heap warp)"
                    f.write(json.dumps(d) + "\n")
print("profiles written")


# ---- index: profiles/r02_README.md ----
def jload(name):
    try:
        return json.loads(open(os.path.join(P, name)).read().strip().split("\n")[-1])
    except Exception:
        return None


def launch_shares():
    """kernel -> (launches, total ms) from the ncu launch list (one row per launch and metric)"""
    path = os.path.join(P, "r02_ncu_launches_bench.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg = {}
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].replace("void ", "").replace("lmb::", "")
        name = name.split("(")[0] if not name.startswith("k_astar<") else name.split(">")[0] + ">"
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


b1, ref = jload("r02_bench_1gpu.json"), jload("r02_bench_reference_arm.json")
multi = {n: jload(f"r02_bench_{n}gpu.json") for n in (2, 4, 8)}
with open(os.path.join(P, "r02_README.md"), "w") as f:
    f.write("# Round-2 profiles — index\n\nEverything here was produced on the pool's B200s by `scripts/r02_measure.sh` (one GPU) or by "
            "the commands quoted\nin the file, and condensed by `scripts/make_r02_profiles.py` (this index included); numbers in "
            "DESIGN.md cite these files.\n\n| file | what |\n|---|---|\n")
    if b1:
        pc = b1.get("parity_check", {})
        f.write(f"| `r02_bench_1gpu.json` | `python bench.py` (config 2, the headline): {b1['ms_per_step']:.1f} ms per step = "
                f"{b1['value']:.0f} agent-updates/s, e2e {b1['e2e']['value']:.0f}/s, roofline frac {b1['roofline']['frac']:.3f}, "
                f"`parity_check` {pc.get('n')} agents ok = {pc.get('ok')}, clocks {b1['clocks']['sm_mhz']:.0f} MHz, throttle reasons "
                f"{b1['clocks']['reasons']} |\n")
    if ref:
        ac = ref.get("all_cores") or {}
        f.write(f"| `r02_bench_reference_arm.json` | `python bench.py --impl reference`: {ref['value']:.1f} agent-updates/s on one core "
                f"(the reference is single-threaded), {ac.get('value', 0):.0f}/s with one oracle process per host core |\n")
    for n in (2, 4, 8):
        d = multi[n]
        if d and b1:
            f.write(f"| `r02_bench_{n}gpu.json` | `torchrun --nproc-per-node {n} bench.py --gpus {n}`: {d['ms_per_step']:.1f} ms per step = "
                    f"{d['value']:.0f} agent-updates/s ({d['value'] / b1['value']:.2f}x the 1-GPU line), library-owned "
                    f"`ncclAllGather`, `allgather_ms` {d.get('phase_ms', {}).get('allgather', 0):.2f} |\n")
    f.write("| `r02_bench_2gpu_config4.json` | `bench.py --gpus 2 --config 4 --agents 250000`: the shared 64-island world, moving agents |\n"
            "| `r02_configs.jsonl` | `bench.py --config 3 / 4 / 5` lines and config 1 (`scripts/run_configs.py 1`), one JSON per line "
            "(`_file` names the run) |\n"
            "| `r02_astar_latency.jsonl` | `scripts/astar_latency.py maze field tiled`: microseconds per explored node of the longest "
            "search per kernel family, 1 ... 2368 queries |\n"
            "| `r02_ncu_dram_field.csv` | `ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
            "gpu__time_duration.sum -k regex:k_astar python scripts/profile_grid_loaded.py`: DRAM bytes of loaded open-field "
            "launches (12 000 antipodal queries on the 256x256 field of 8 m quads, 4 736 slots; 338.7 M explored nodes per launch) |\n"
            "| `r02_layout_latency.jsonl` | `scripts/layout_latency.py`: dense against hashed `best_estimates` on the 1 M-polygon "
            "tiled world, warp-pair kernel, 1 and 148 queries (what the dense side arena of small batches buys) |\n"
            "| `r02_ncu_launches_bench.csv` | `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` of `bench.py --steps 2 "
            "--warmup 1`: every kernel launch of the bench (cold-cache, serialised: compare shares) |\n")
    try:
        vals = {}
        for r in csv.reader(l for l in open(os.path.join(P, "r02_ncu_dram_bench.csv")) if l.startswith('"')):
            if r[0] != "ID":
                vals[r[-3]] = float(r[-1].replace(",", ""))
        rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
        f.write(f"| `r02_ncu_dram_bench.csv` | DRAM bytes of the bench's own 100 000-query `k_astar` launch (ncu application replay): "
                f"{rd / 1e9:.1f} GB read + {wr / 1e9:.1f} GB written = {(rd + wr) / 6.479e9:.1f} B per explored node |\n")
    except Exception as e:  # noqa: BLE001
        f.write(f"| `r02_ncu_dram_bench.csv` | DRAM bytes of the bench-size `k_astar` launch ({e}) |\n")
    f.write("| `r02_ncu_astar_loaded.md`, `_raw.csv` | `ncu --set full` of a fully loaded `k_astar<16,40,5,quad,forest,SIMPLE>` launch "
            "on a capturable arena, with the per-line stall table |\n"
            "| `r02_ncu_astar_single.md` | `ncu --set full` of ONE search: warp kernel before / after `cp.async` staging, thread kernel, "
            "open field; the false scoreboard wait; the final warp-pair kernel |\n"
            "| `r02_ncu_avoid_follow_cfg3.md` | `ncu --set full` of `k_avoidance` and `k_follow` at config 3 (1 M agents) |\n")
    try:
        agg = launch_shares()
        total = sum(v[1] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])
        f.write("\nShare of the bench step by kernel (from `r02_ncu_launches_bench.csv`; the un-profiled run reports `phase_ms.astar` = "
                f"{b1['phase_ms']['astar']:.1f} of {b1['ms_per_step']:.1f} ms = {100 * b1['phase_ms']['astar'] / b1['ms_per_step']:.1f} %):\n\n"
                "| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for name, (cnt, ms) in top[:8]:
            f.write(f"| `{name}` | {cnt} | {ms:.1f} | {100 * ms / total:.2f} % |\n")
        rest = top[8:]
        f.write(f"| everything else ({sum(v[0] for _, v in rest)} launches of {len(rest)} kernels) | | {sum(v[1] for _, v in rest):.1f} | "
                f"{100 * sum(v[1] for _, v in rest) / total:.2f} % |\n")
    except Exception as e:  # noqa: BLE001
        f.write(f"\n(launch shares unavailable: {e})\n")
print("index written")
