#!/usr/bin/env python
"""Turns the ncu reports that a gpurun call left in gpurun_out/ into the text/CSV summaries committed under profiles/.

    python tools/summarize_profiles.py [round_tag]        (default r01)

Needs the ncu CLI (reads .ncu-rep files; no GPU).  Inputs (all optional), written by tools/final_suite.sh <tag>:
gpurun_out/launches_<tag>.csv (launch list from `ncu --metrics gpu__time_duration.sum`) and
gpurun_out/{scan,gemm,elem,ssd}_<tag>.ncu-rep (`ncu --set full` captures).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

COMMON = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
          'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
          'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']


def raw_rows(rep):
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return None
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows if len(rows) > 2 else None


def dump(rep, out, title, stalls=False):
    rows = raw_rows(rep)
    if rows is None:
        return None
    hdr, units = rows[0], rows[1]
    first = None
    with open(os.path.join(PROF, out), "w") as f:
        f.write(title + "\n")
        for r in rows[2:]:
            if len(r) < 10:
                continue
            kn = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
            f.write(f"\n== {kn}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            vals = {}
            for k in COMMON:
                if k in hdr:
                    i = hdr.index(k)
                    vals[k] = (r[i], units[i])
                    f.write(f"  {k:100s} {r[i]:>18s} {units[i]}\n")
            if first is None:
                first = vals
            if stalls:
                st = []
                for i, h in enumerate(hdr):
                    if 'smsp__pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
                        try:
                            st.append((float(r[i].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
                        except ValueError:
                            pass
                tot = sum(v for v, _ in st) or 1.0
                f.write("  warp stall sampling (share of samples):\n")
                for v, h in sorted(st, reverse=True)[:10]:
                    f.write(f"    {100 * v / tot:5.1f}%  {h}\n")
    return first


def launches():
    p = os.path.join(OUT, f"launches_{TAG}.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'Block Size'))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
        key = (r[ki].split('(')[0].replace('void ', ''), r[gi], r[bi])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)",
             "# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-clocks",
             "# window: 300 consecutive launches inside the forward pass of PlantCaduceus_l32, B=256 x 512 bp",
             f"# total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches", "",
             "share_pct,total_ms,launches,avg_ms,kernel,grid,block"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{100 * v[1] / tot:.2f},{v[1]:.3f},{v[0]},{v[1] / v[0]:.4f},{k[0]},\"{k[1]}\",\"{k[2]}\"")
    open(os.path.join(PROF, f"{TAG}_launches_summary.csv"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def main():
    os.makedirs(PROF, exist_ok=True)
    launches()
    scan = dump(f"scan_{TAG}.ncu-rep", f"{TAG}_scan_ncu_summary.txt",
                "ncu --set full --clock-control none --import-source on -k regex:biscan -s 3 -c 1  python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-clocks\n"
                "pcad::biscan_kernel<__nv_bfloat16, false, kScanBcF32 = 2>: fp32 B|C rows by TMA, one block barrier per chunk  (PlantCaduceus_l32, B=256 x 512 bp: S=512 sequences, E=2048, one launch per layer)\n"
                "algorithmic bytes per launch 6.543 GB (817.9 MB/window / 32 layers x 256 windows)", stalls=True)
    if scan:
        rd, wr = scan['dram__bytes_read.sum'], scan['dram__bytes_write.sum']
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        traffic = float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]]
        json.dump({"l32": traffic, "source": f"profiles/{TAG}_scan_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch, B=256)"},
                  open(os.path.join(PROF, "scan_traffic.json"), "w"))
        print(open(os.path.join(PROF, f"{TAG}_scan_ncu_summary.txt")).read())
    dump(f"gemm_{TAG}.ncu-rep", f"{TAG}_gemm_ncu_summary.txt",
         "ncu --set full --clock-control none -k regex:gemm_bf16 -s 8 -c 6 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-clocks\n"
         "PlantCaduceus_l32, B=256 (T = 262144 strand-tokens). Launch order within a layer: in_proj<256, row-scale epilogue> [T,1024]x[4096,1024]; "
         "x_proj<96> [T,2048]x[96,2048]; dt_proj<256> [T,64]x[2048,64]; (x_proj, dt_proj again for the reverse direction); "
         "out_proj<256, residual epilogue> [T,2048]x[1024,2048].  UTCHMMA = tcgen05.mma.")
    dump(f"elem_{TAG}.ncu-rep", f"{TAG}_conv_norm_ncu_summary.txt",
         "ncu --set full --clock-control none -k regex:'conv_silu|add_rmsnorm' python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-clocks   (PlantCaduceus_l32, B=256)")


    dump(f"ssd_{TAG}.ncu-rep", f"{TAG}_ssd_ncu_summary.txt",
         "ncu --set full --clock-control none --import-source on -k regex:'ssd_chunk_tc_kernel|gated_norm_sum' -s 4 -c 2 "
         "python bench.py --workload long --model cad2-small --batch 8 --steps 1 --warmup 1 --cpu-sample 0 --no-clocks\n"
         "PlantCAD2-Small (Mamba-2: d 768, 24 heads of 64, d_state 64), B = 8 x 8192 bp: S = 16 sequences, T = 131072 strand-tokens;\n"
         "ssd_chunk_tc_kernel: grid (12 head pairs, 16 sequences, 2 directions), 64 chunks of 128 positions per CTA, one launch per layer.\n"
         "algorithmic bytes per launch: 2 directions x T x (x 3 KB + B,C 256 B read, y 3 KB written) = 1.68 GB", stalls=True)


if __name__ == "__main__":
    main()
