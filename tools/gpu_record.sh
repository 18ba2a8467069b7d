#!/bin/bash
# the round's record run: the default bench line and the reference arm as the driver runs them (wall time of each),
# then the ncu launch list of the quick bench command and one --set full capture of the C2 step's three kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-rec}
S=$(date +%s); timeout 900 python bench.py > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench.py rc=$? wall=$(( $(date +%s) - S ))s"
tail -c 600 $O/${T}_bench.err
S=$(date +%s); timeout 900 python bench.py --impl reference > $O/${T}_ref.log 2> $O/${T}_ref.err; echo "bench.py --impl reference rc=$? wall=$(( $(date +%s) - S ))s"
python - $O/${T}_bench.log $O/${T}_ref.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
r=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][-1])
print("ours: value %.4g %s ms_per_step %.3f e2e %.4g (%.3f ms) frac %.3f whole %.3f / %.3f launches %s clocks %s" % (
    d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'],
    d['roofline']['whole_step_frac'], d['roofline']['whole_step_frac_e2e'], d.get('gpu_launches'), d.get('clocks')))
print("cpu_baseline", d.get('cpu_baseline'))
print("keys", sorted(d.keys()))
print("ref: value %.4g %s e2e %s cpu_baseline %s" % (r['value'], r['unit'], r.get('e2e'), r.get('cpu_baseline')))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2z_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mf_owner_kernel|owner_schedule_tab|owner_setup" -c 3 -o $O/r2z_owner -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $O/${T}_ncu_owner.log 2>&1
ls -la $O/r2z_*
