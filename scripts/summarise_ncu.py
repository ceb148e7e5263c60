"""Turn ncu outputs into the small, reviewable files kept under profiles/.

  launch list  (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...):
      python scripts/summarise_ncu.py launches X.csv profiles/NAME          -> NAME_launches.csv (id,kernel,grid,block,ns)
                                                                               NAME_by_kernel.tsv (share of GPU time)
  full capture (ncu --set full -o X ...; exported with `ncu -i X.ncu-rep --page raw --csv > X_raw.csv`):
      python scripts/summarise_ncu.py full X_raw.csv profiles/NAME.tsv      -> the metrics DESIGN.md / bench.py cite
"""
import collections
import csv
import re
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"\(unnamed\)::|<unnamed>::|dfw::|void ", "", name)
    return name.split("(")[0].strip()


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ik, ig, ib, iv = (hdr.index(c) for c in ("Kernel Name", "Grid Size", "Block Size", "Metric Value"))
    agg, cnt = collections.Counter(), collections.Counter()
    with open(dst + "_launches.csv", "w") as f:
        f.write("id,kernel,grid,block,gpu_time_ns\n")
        for i, r in enumerate(rows[1:]):
            k = short(r[ik])
            ns = float(r[iv].replace(",", ""))
            agg[k] += ns; cnt[k] += 1
            f.write(f'{i},"{k}","{r[ig]}","{r[ib]}",{ns:.0f}\n')
    tot = sum(agg.values())
    with open(dst + "_by_kernel.tsv", "w") as f:
        f.write(f"# per-kernel share of GPU time over {len(rows) - 1} profiled launches (serialised, cold-cache: shares, not absolutes)\n")
        f.write("kernel\tlaunches\tms\tshare\n")
        for k, v in agg.most_common():
            f.write(f"{k}\t{cnt[k]}\t{v / 1e6:.3f}\t{v / tot:.4f}\n")


def full(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w") as f:
        f.write("kernel\t" + "\t".join(f"{hdr[c]} [{units[c]}]" for c in cols) + "\n")
        for r in rows[2:]:
            f.write(short(r[ik]) + "\t" + "\t".join(r[c] for c in cols) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
