# tools/profile_summarise.sh <tag>   (here, no GPU): gpurun_out/prof_<tag>_{bwd,fwd}.ncu-rep + launches_<tag>.csv -> profiles/
tag=${1:-r02}
cd "$(dirname "$0")/.."
for k in bwd fwd; do
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv | python tools/ncu_summary.py profiles/ncu_${tag}_${k}_summary.csv
done
cp gpurun_out/launches_$tag.csv profiles/ncu_${tag}_launches.csv
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
def traffic(path):
    rows = {r[0]: r[2] for r in csv.reader(open(path)) if len(r) == 3}
    unit = {r[0]: r[1] for r in csv.reader(open(path)) if len(r) == 3}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    return sum(float(rows[k]) * scale[unit[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
old = json.load(open("profiles/roofline_traffic.json"))
old["note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from ncu --set full captures "
               f"(cfg3: profiles/ncu_{tag}_{{fwd,bwd}}_summary.csv of the shipped library; cfg2: profiles/ncu_r01b_tf32_summary.csv)")
old["cfg3"] = {"fwd": traffic(f"profiles/ncu_{tag}_fwd_summary.csv"), "bwd": traffic(f"profiles/ncu_{tag}_bwd_summary.csv")}
json.dump(old, open("profiles/roofline_traffic.json", "w"), indent=1)
print(old["cfg3"])
PY
python tools/sass_listing.py
