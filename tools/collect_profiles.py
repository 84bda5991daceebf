#!/usr/bin/env python
"""Turns the outputs of tools/final_run.sh (gpurun_out/r2_*) into the tracked files under profiles/ (round 2).
usage: python tools/collect_profiles.py"""
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
METRICS = ("gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum lts__t_sector_hit_rate.pct "
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__m_xbar2l1tex_read_bytes.sum launch__registers_per_thread "
           "launch__grid_size launch__block_size sm__warps_active.avg.pct_of_peak_sustained_active "
           "smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum "
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active "
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active smsp__pcsamp_warps_issue_stalled "
           "lts__t_sectors_srcunit_tex_op_red.sum").split()


def last_json_line(path):
    return [l for l in open(path) if l.startswith("{")][-1]


def sh(cmd, stdin=None):
    return subprocess.run(cmd, shell=True, input=stdin, capture_output=True, text=True).stdout


def main():
    for a, b in {"r2_bench": "r02_bench_line.json", "r2_bench_ref": "r02_bench_line_reference_arm.json",
                 "r2_bench_fcos": "r02_bench_line_fcos.json", "r2_bench_d3": "r02_bench_line_1gpu_3domains.json"}.items():
        (PROF / b).write_text(last_json_line(OUT / f"{a}.log"))
    for a, b in {"r2_opbench_roi_608": "r02_op_sweep_roi_608x1024", "r2_opbench_roi_800": "r02_op_sweep_roi_800x1344",
                 "r2_opbench_other": "r02_op_sweep_other"}.items():
        shutil.copy(OUT / f"{a}.json", PROF / f"{b}.json")
    bwd_us = [l for l in open(OUT / "r2_bwd_check.log") if l.startswith("algo 4")][-1].split()[2]
    traffic = {}
    for rep, name, key, cmd, note in (
            ("r2_own_bwd", "r02_ncu_full_roi_align_own_bwd.txt", "msroi_align_bwd",
             "-k regex:own_bwd -s 2 -c 1 python tools/bwd_check.py --algos 4 --iters 2",
             f"plain run of the same command: {bwd_us} us for the whole C-ABI call incl. the plan and bin launches, CUDA events, L2 flushed"),
            ("r2_fwd", "r02_ncu_full_roi_align_fwd.txt", "msroi_align_fwd",
             "-k regex:msroi_fwd_tma -s 2 -c 1 python tools/profile_roi.py", "plain runs: tools/op_bench.py rows of r02_op_sweep_roi_608x1024.json")):
        raw = sh(f"ncu -i {OUT / (rep + '.ncu-rep')} --page raw --csv 2>/dev/null")
        txt = sh(f"{sys.executable} {ROOT / 'tools' / 'ncu_raw.py'} " + " ".join(METRICS), stdin=raw)
        txt = "\n".join(l for l in txt.splitlines() if "not_issued" not in l)
        (PROF / name).write_text(f"# ncu --set full --clock-control none --import-source on {cmd}\n"
                                 f"# (B=8, 512 RoIs/img, C=256, 4 levels of a 608x1024 image, fp32 NHWC; algorithmic bytes 629.0 MB; {note})\n" + txt + "\n")
        rd = wr = 0.0
        for l in txt.splitlines():
            p = l.split()
            if p and p[0] == "dram__bytes_read.sum": rd = float(p[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[p[2]]
            if p and p[0] == "dram__bytes_write.sum": wr = float(p[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[p[2]]
        traffic[key] = int(rd + wr)
        src = sh(f"ncu -i {OUT / (rep + '.ncu-rep')} --page source --csv --print-source cuda,sass 2>/dev/null")
        tmp = OUT / f"{rep}_src.csv"
        tmp.write_text(src)
        lines = sh(f"{sys.executable} {ROOT / 'tools' / 'ncu_lines.py'} {tmp} 40")
        (PROF / name.replace("ncu_full", "ncu_source")).write_text(
            f"# stall samples per CUDA source line (same capture as {name}; tools/ncu_lines.py)\n" + lines)
    traffic["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (B=8, 512 RoIs/img, C=256, 608x1024, "
                           "fp32 NHWC): forward from r02_ncu_full_roi_align_fwd.txt, backward from r02_ncu_full_roi_align_own_bwd.txt "
                           "(main kernel; the plan and bin launches move ~11 MB more)")
    (PROF / "roofline_traffic.json").write_text(json.dumps(traffic, indent=1))
    ll = sh(f"{sys.executable} {ROOT / 'tools' / 'ncu_launches.py'} {OUT / 'r2_launches_bench.csv'} 45")
    (PROF / "r02_ncu_launches_bench_partial.txt").write_text(
        "# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1800 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e\n"
        "# (timed region of the bench only: the first 1 800 launches; a number printed by the run under ncu is not a bench value)\n" + ll)
    print("profiles/ updated;", traffic)


if __name__ == "__main__":
    main()
