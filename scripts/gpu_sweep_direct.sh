cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # batch direct cluster
  LL_ASSOC_DIRECT=$2 LL_LM_CLUSTER=$3 timeout 300 python bench.py --batch $1 --steps 20 --warmup 3 --no-cpu --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch $1 direct $2 cluster $3: value %.0f ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
}
for b in 4 16; do for d in 0 1000; do for cl in 1 8; do run $b $d $cl; done; done; done
for b in 32 64; do for d in 0 1000; do for cl in 1 2; do run $b $d $cl; done; done; done
for b in 128 256; do for d in 0 1000; do run $b $d 1; done; done
