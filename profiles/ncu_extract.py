"""Export the per-launch raw metrics of an .ncu-rep as a small CSV that is committed and that bench.py reads for
`roofline.traffic` (no literal in the bench):

    python profiles/ncu_extract.py gpurun_out/xattn.ncu-rep profiles/xattn_fused_ncu_raw.csv [kernel-substring]

columns: kernel, duration_ns, dram_bytes_read, dram_bytes_write, tensor_pipe_pct, dram_throughput_pct, lts_throughput_pct,
sm_warps_active_pct, registers
(`ncu -i <rep> --page raw --csv`; metric names per /opt/skills/guides/B200_PROFILING.md)"""
import csv
import subprocess
import sys

rep, dst = sys.argv[1], sys.argv[2]
sub = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def col(*names):
    for n in names:
        if n in hdr:
            return hdr.index(n)
    return None


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_ns(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(u, 1)


c = dict(name=col("Kernel Name"), dur=col("gpu__time_duration.sum"), rd=col("dram__bytes_read.sum"),
         wr=col("dram__bytes_write.sum"),
         tp=col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
         dt=col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         lt=col("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         wa=col("sm__warps_active.avg.pct_of_peak_sustained_active"), rg=col("launch__registers_per_thread"))
with open(dst, "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow(["kernel", "duration_ns", "dram_bytes_read", "dram_bytes_write", "tensor_pipe_pct",
                 "dram_throughput_pct", "lts_throughput_pct", "sm_warps_active_pct", "registers"])
    for r in rows[2:]:
        if not r or sub not in r[c["name"]]:
            continue
        g = lambda k: (r[c[k]] if c[k] is not None else "")
        wr.writerow([r[c["name"]][:120], to_ns(g("dur"), units[c["dur"]]), to_bytes(g("rd"), units[c["rd"]]),
                     to_bytes(g("wr"), units[c["wr"]]), g("tp"), g("dt"), g("lt"), g("wa"), g("rg")])
print("wrote", dst)
