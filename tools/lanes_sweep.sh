# sweep of the lanes-per-row choice of the transpose pass (x-phase) against gather locality
run() { HPRLP_LANES_AT=$2 SYNTH_BLOCK_RUN=$3 timeout 300 python bench.py --workload $1 --steps 5 --warmup 3 --no-batched --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$1 run=$3 AT=$2', round(d['value'],1), round(d['roofline']['x_phase']['ms'],4), round(d['roofline']['y_phase']['ms'],4))"; }
for g in 1 2 4; do run c3band $g 8; done
for r in 2 4; do for g in 1 2 4; do run c3block $g $r; done; done
for g in 1 2 4; do HPRLP_LANES_A=$g run c2 $g 8; done
