#!/bin/bash
# Produces the round's judged evidence under gpurun_out/ (copied into profiles/ afterwards):
#   bench line (default command), ncu launch list of the same command, one --set full capture of the
#   dominant kernel at the full 1M-window size (DRAM traffic per launch) with the details page.
R=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench.err
tail -c 600 gpurun_out/${R}_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-compute-roofline > gpurun_out/${R}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_kernel -s 1 -c 1 -o gpurun_out/${R}_prof -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-compute-roofline > gpurun_out/${R}_ncu.log 2>&1
ncu -i gpurun_out/${R}_prof.ncu-rep --page details > gpurun_out/${R}_poa_kernel_Tc_1M_details.txt 2>/dev/null
tail -2 gpurun_out/${R}_ncu.log
ncu -i gpurun_out/${R}_prof.ncu-rep --page raw --csv > gpurun_out/${R}_raw.csv 2>/dev/null
python - "$R" <<'PY'
import csv, json, sys
R = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/{R}_raw.csv")))
h, v = rows[0], rows[2]
d = dict(zip(h, v))
rd, wr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
u = dict(zip(h, rows[1]))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rd *= scale.get(u["dram__bytes_read.sum"], 1); wr *= scale.get(u["dram__bytes_write.sum"], 1)
json.dump({"windows": 1000000, "arms": 30, "length": 120, "kernel": d.get("Kernel Name", "poa_kernel"),
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "source": f"ncu --set full, profiles/{R}_poa_kernel_Tc_1M_details.txt"}, open(f"gpurun_out/{R}_traffic.json", "w"), indent=1)
print(open(f"gpurun_out/{R}_traffic.json").read())
PY
