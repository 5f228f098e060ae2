#!/bin/bash
# round-2 session g: single-barrier Sinkhorn (fixed-point atomic column sums): parity + A/B, persisting-L2 experiment
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== sinkhorn-related tests" | tee gpurun_out/g_p1.log
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 -k "sinkhorn or c2_full or coupling_stream or trajectory or fast_draw or baseline_config or zero_mass or reference_fm" >> gpurun_out/g_p1.log 2>&1
echo "rc=$?" >> gpurun_out/g_p1.log; tail -12 gpurun_out/g_p1.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-ode --no-cpu-baseline --no-extra > gpurun_out/g_$label.json 2> gpurun_out/g_$label.err
  python -c "import json;d=json.load(open('gpurun_out/g_$label.json'));print('$label value',round(d['value'],2),'solve',round(d['roofline']['stage_ms']['solve'],3),'parity',{k:(round(v,9) if isinstance(v,float) else v) for k,v in d['parity'].items()})"
}
run atomic1 CFM_SK_ATOMIC=1
run atomic0 CFM_SK_ATOMIC=0
run atomic1b CFM_SK_ATOMIC=1
run atomic0b CFM_SK_ATOMIC=0
run a1_persist100_l2_15 CFM_SK_ATOMIC=1 CFM_SK_PERSIST_MB=100 CFM_SK_L2=0.15
run a1_persist100_l2_30 CFM_SK_ATOMIC=1 CFM_SK_PERSIST_MB=100 CFM_SK_L2=0.30
run a1_persist100_l2_40 CFM_SK_ATOMIC=1 CFM_SK_PERSIST_MB=100 CFM_SK_L2=0.40
run a1_persist64_l2_25 CFM_SK_ATOMIC=1 CFM_SK_PERSIST_MB=64 CFM_SK_L2=0.25
run a1_l2_10 CFM_SK_ATOMIC=1 CFM_SK_L2=0.10
run a1_l2_20 CFM_SK_ATOMIC=1 CFM_SK_L2=0.20
