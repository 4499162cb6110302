#!/bin/bash
# The reference's multi-GPU command lines (train.py, inference.py) under torchrun on N GPUs of one box:
#   bash benchmarks/multi_gpu_user_path.sh [N]
# writes a synthetic nerf_synthetic-shaped directory, trains 6 epochs ray-sharded (NCCL gradient all-reduce),
# renders the orbit with frames sharded over the ranks, and prints rank 0's log.csv.
set -e
N=${1:-2}
T=${TMPDIR:-/tmp}/knerf_mg
rm -rf "$T"; mkdir -p "$T"
python -c "from keras_nerf_b200.data.synthetic import write_nerf_synthetic_like as w; w('$T/scene', image_wh=128, n_train=16, n_val=$N, n_test=$((4*N)))"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$RUN train.py --name ball --data_dir "$T/scene" --img_wh 64 --ray_chunks 4096 --num_epochs 6 \
     --log_dir "$T/logs" --model_dirs "$T/model" --log_freq 2 > "$T/train.log" 2>&1 || { grep -v Warning "$T/train.log" | grep -B2 -A12 Traceback | head -60; exit 1; }
tail -2 "$T/train.log"
cat "$T/logs/ball/log.csv" | cut -d, -f1,2,3,5,6,8,11
$RUN inference.py --model_dirs "$T/model/ball" --img_wh 64 --output_freq 30 --output_dir "$T/out" > "$T/inf.log" 2>&1 || { tail -40 "$T/inf.log"; exit 1; }
ls -l "$T/out"
