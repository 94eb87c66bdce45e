cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "vote_partial or deskew" 2>&1 | tail -6
for K in 1 3 6; do
timeout 900 python bench.py --config 4 --stream-scans 1280 --lanes 64 --overlap $K > gpurun_out/r02_config4_n1_k$K.json 2> gpurun_out/cfg4.err; tail -2 gpurun_out/cfg4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_config4_n1_k$K.json')); print($K, d['value'], d['e2e']['value'], d['exchange'], d['deviation_vs_unsegmented_chain'])"
done
