"""Turns the csv exports of scratch/r2_final.sh (gpurun_out/r2z_*) into the committed summaries under profiles/."""
import csv, gzip, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_red.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_active.avg',
        'sm__cycles_elapsed.avg', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def raw_summary(name, command, note):
    rows = list(csv.reader(open(os.path.join(G, f"r2z_{name}_raw.csv"))))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        kernels.append({w: (r[hdr.index(w)] + (" " + units[hdr.index(w)] if units[hdr.index(w)] else "")) for w in WANT if w in hdr})
    out = {"command": command, "captured_at_commit": commit, "note": note, "kernels": kernels}
    json.dump(out, open(os.path.join(P, f"r2z_{name}_ncu_summary.json"), "w"), indent=1)
    return rows, hdr


def sass_hist(name):
    txt = gzip.open(os.path.join(G, f"r2z_{name}_src.csv.gz"), "rt").read()
    rows = list(csv.reader(io.StringIO(txt)))
    hists, lines = collections.OrderedDict(), {}
    kern, hdr, fileName = None, None, None

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fileName = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            kern = r[1]
            hists.setdefault(kern, collections.Counter()); lines.setdefault(kern, {})
            continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or kern is None:
            continue
        iI = hdr.index("Instructions Executed")
        if r[0].isdigit():
            key = (fileName, int(r[0]))
            old = lines[kern].get(key, ("", 0))
            lines[kern][key] = (r[1], old[1] + num(r[iI]))
        elif r[0] == "" and len(r) > iI and r[2].startswith("0x"):
            ops = [o for o in r[3].split() if not o.startswith("@")]
            if ops:
                hists[kern][ops[0].split(".")[0]] += num(r[iI])
    out = []
    for kern, hist in hists.items():
        t = sum(hist.values())
        if not t:
            continue
        out.append(f"== {kern}\n   executed warp instructions (all captured launches of this function): {t}")
        out.append("   opcode histogram (share of executed warp instructions): " +
                   ", ".join(f"{k} {v / t * 100:.1f}%" for k, v in hist.most_common(18)))
        out.append("   hottest source lines:")
        for (f, ln), (src, n) in sorted(lines[kern].items(), key=lambda kv: -kv[1][1])[:12]:
            out.append(f"     {n / t * 100:5.1f}%  {f}:{ln}  {src.strip()[:110]}")
    return "\n".join(out)


if __name__ == "__main__":
    caps = {
        "headline": ("ncu --set full --clock-control none --import-source on -k regex:'sep2d_fwd_kernel|sep2d_adj_stream' -s 8 -c 2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-multi",
                     "config #2 (Lanczos-4, Float32 4096^2 -> 8192^2): one forward and one adjoint launch"),
        "3d_fwd": ("ncu --set full ... -k regex:sep2d_fwd -s 1 -c 1 python bench_configs.py --configs 4", "config #4 (1024^3, fused 3-D forward, D3 = true)"),
        "3d_adj": ("ncu --set full ... -k regex:sep2d_adj_stream -s 1 -c 1 python bench_configs.py --configs 4", "config #4 (1024^3, fused 3-D adjoint, D3 = true)"),
        "multi": ("ncu --set full ... -k regex:sep2d_adj_multi -s 2 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 512",
                  "config #5 (512 frames of 2048^2 into one model, one launch)"),
    }
    sass = []
    traffic = {"captured_at": commit}
    for name, (cmd, note) in caps.items():
        rows, hdr = raw_summary(name, cmd, note)
        if name == "headline":
            for r in rows[2:]:
                k = r[hdr.index("Kernel Name")]
                rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
                unit = rows[1][hdr.index("dram__bytes_read.sum")]
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]
                traffic["sep2d_fwd_kernel" if "fwd" in k else "sep2d_adj_stream_kernel"] = int((rd + wr) * scale)
        sass.append(f"#### {name}: {note}\n" + sass_hist(name))
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    open(os.path.join(P, "r2z_sass_histograms.txt"), "w").write(
        "Executed-SASS histograms and hottest source lines of the ncu captures of this round (scratch/make_profiles.py from the\n"
        f"`--page source --print-source cuda,sass` export; commit {commit}).\n\n" + "\n\n".join(sass) + "\n")
    for f in ("bench.json", "launches.csv", "configs.log", "spectrum.log"):
        dst = {"bench.json": "r2z_bench_line_1gpu.json", "launches.csv": "r2z_launch_list_gpu_time.csv", "configs.log": "r2z_other_configs.jsonl",
               "spectrum.log": "r2z_rate_spectrum.txt"}[f]
        open(os.path.join(P, dst), "w").write(open(os.path.join(G, "r2z_" + f)).read())
    print(open(os.path.join(P, "traffic.json")).read())
