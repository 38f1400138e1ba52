#!/usr/bin/env bash
# One GPU call that produces an ncu capture for every kernel family of libkge_b200 (north_star: "each kernel ships
# with an ncu capture").  Run on the B200 box from the repo root, ONE GPU (ncu replays each kernel ~40x):
#   gpurun --timeout 900 -- 'bash scripts/ncu_all.sh r2'
# Outputs (small CSV/JSON summaries, no .ncu-rep) land in gpurun_out/; copy what should be judged into profiles/.
set -uo pipefail
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"

cap() {  # cap <name> <kernel regex> <skip> <count> <cmd...>
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  timeout 300 $NCU --set full --import-source on -k "regex:$regex" -s "$skip" -c "$count" -f -o "$out/${tag}_$name" "$@" > "$out/${tag}_$name.stdout" 2>&1
  ncu -i "$out/${tag}_$name.ncu-rep" --page raw --csv > "$out/${tag}_${name}_raw.csv" 2>/dev/null
  python scripts/ncu_summary.py full "$out/${tag}_${name}_raw.csv" "$out/${tag}_${name}_ncu_full_summary.json" \
      "ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count $*"
  rm -f "$out/${tag}_$name.ncu-rep" "$out/${tag}_${name}_raw.csv"
}

# launch list of the bench command (kernel shares of a step)
timeout 300 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file "$out/${tag}_launches_bench.csv" \
    python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
python scripts/ncu_summary.py launches "$out/${tag}_launches_bench.csv" "$out/${tag}_launch_list_summary.csv" \
    "ncu --metrics gpu__time_duration.sum --clock-control none -c 300 python bench.py --steps 5 --warmup 3 --no-cpu"

# training kernels: resident ComplEx (cfg2), resident DistMult NIT=4 (cfg3), grouped RotatE (cfg4), HBM-resident table (big)
cap train_cfg2 'kge_train_kernel|kge_optim_kernel' 8 2 python bench.py --steps 5 --warmup 3 --no-cpu
cap train_cfg3 kge_train_kernel 2 1 python scripts/kbench.py one cfg3 red_v4 0
cap train_cfg4 kge_train_kernel 2 1 python scripts/kbench.py one cfg4 red_v4 0
cap train_big  kge_train_kernel 2 1 python scripts/kbench.py one big red_v4 0
# ranking kernels: FFMA2 dot kernel (ComplEx), generic tile kernel (TransE, RotatE), prepare / filter / finalize
cap rank_dot    'kge_rank_dot_kernel|kge_rank_q|kge_rank_finalize' 0 4 python scripts/rbench.py one ComplEx 200 14505 1024
cap rank_transe kge_rank_tile_kernel 0 1 python scripts/rbench.py one TransE 400 14505 1024
cap rank_rotate kge_rank_tile_kernel 0 1 python scripts/rbench.py one RotatE 200 14505 1024
ls -la $out | tail -20
