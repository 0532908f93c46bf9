# Per-stage cost model of the A-resident PatchSelecter kernels.  Needs a profile build of the library:
#   NVCC_EXTRA="-DQAT_PROFILE_BUILD" python qa-tiger_b200/build.py --force
# flags (option gemm_attn_ts_flags, GemmAttnP::flags): 4 no attention math, 8 no drain, 16 no MMAs, 32 no TMA loads; e.g.
#   FLAGS="0 12 16 28 44" bash tests/ncu_flags.sh     (everything / no epilogue / no MMAs / loads only / MMAs only)
# Prints per kernel: gpu__time_duration (ns), tensor pipe active %, DRAM bytes read -- min and median over the launches.
for f in ${FLAGS:-0 12 16 28 44}; do
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum --clock-control none -k regex:gemm_attn --csv --log-file gpurun_out/flags_$f.csv python tests/prof_patch_attn.py 256 60 $f > /dev/null 2>&1
  echo "flags=$f"; python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/flags_$f.csv")) if len(r)>10]
hdr=rows[0]
ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value")
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]:
    d[(r[ik][:40], r[im])].append(float(r[iv].replace(",","")))
for k,v in sorted(d.items()):
    print(k, "n=%d"%len(v), "min=%.1f med=%.1f"%(min(v), sorted(v)[len(v)//2]))
P
done
