"""Text summaries of the round's `ncu --set full` captures (tools/profile_round.sh) for profiles/.

  python tools/ncu_summaries.py gpurun_out/r01_v10 profiles/r01 v10
"""
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return {h: i for i, h in enumerate(hdr)}, units, data


def num(r, idx, k):
    try:
        return float(r[idx[k]].replace(",", ""))
    except Exception:
        return float("nan")


def conv(rep, dst):
    idx, units, data = raw(rep)
    L = ["ncu --set full --clock-control none: tensor-core launches of one perceive step (B=4) in launch order: the temporal",
         "model (entry conv + block_fused per TemporalBlock, aspp_fused x2 = DeepLab head), then the first decoder convs.  Cold cache.",
         "PAIR=1: tcgen05.mma.cta_group::2 tiling; STACK=1: stacked [W_hi; W_lo] operand.  dram = dram__bytes_read/write.sum,",
         "L2->SM = lts__t_sectors_srcunit_tex_op_read.sum * 32 B.",
         f"{'#':>2} {'kernel':24s} {'grid':>5} {'us':>8} {'dramRd MB':>10} {'dramWr MB':>10} {'L2->SM GB':>10} {'L2->SM TB/s':>12} {'regs':>5}"]
    trd = twr = tt = 0.0
    for i, r in enumerate(data):
        name = r[idx["Kernel Name"]]
        short = name.split("conv_igemm_kernel")[1].split("(")[0] if "conv_igemm" in name else \
            ("  aspp_fused" if "aspp_fused" in name else "  block_fused" if "block_fused" in name else name[:20])
        t = num(r, idx, "gpu__time_duration.sum")
        rd, wr = num(r, idx, "dram__bytes_read.sum"), num(r, idx, "dram__bytes_write.sum")
        l2 = num(r, idx, "lts__t_sectors_srcunit_tex_op_read.sum") * 32 / 1e9
        L.append(f"{i:2d} {'conv_igemm' if 'conv_igemm' in name else '          '}{short:14s} {r[idx['launch__grid_size']]:>5} {t:8.1f} {rd:10.1f} {wr:10.1f} {l2:10.2f} "
                 f"{l2 / t * 1e3:12.2f} {r[idx['launch__registers_per_thread']]:>5}")
        trd += rd; twr += wr; tt += t
    L.append(f"units: {units[idx['gpu__time_duration.sum']]}, {units[idx['dram__bytes_read.sum']]}")
    L.append(f"sum: {tt:.1f} us, dram read {trd:.0f} MB, write {twr:.0f} MB")
    open(dst, "w").write("\n".join(L) + "\n")
    print("\n".join(L))


def liftsplat(rep, dst):
    idx, units, data = raw(rep)
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
    L = ["ncu --set full --clock-control none: the two lift-splat kernels of one perceive step (B=4 samples, 72 camera "
         "frames); cold cache"]
    for r in data:
        L.append(r[idx["Kernel Name"]][:90])
        for k in keys:
            if k in idx:
                L.append(f"   {k:88s} {r[idx[k]]:>16s} {units[idx[k]]}")
    open(dst, "w").write("\n".join(L) + "\n")
    print("\n".join(L))


if __name__ == "__main__":
    src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    conv(src + "_conv_full.ncu-rep", f"{dst}_ncu_conv_{tag}_summary.txt")
    liftsplat(src + "_liftsplat_full.ncu-rep", f"{dst}_ncu_liftsplat_{tag}_summary.txt")
