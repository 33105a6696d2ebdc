#!/bin/bash
# Experiment grid driver — same interface as the reference's run.sh (reference run.sh:1-52):
#   ./run.sh WORLD_SIZE BATCH_SIZE EPOCH_SIZE LEARNING_RATE GPUSET [DE]
# runs {dbs on, off} x {cifar10, cifar100} x {resnet, densenet, googlenet, regnet} with -ocp true.
if [[ $# -ne 5 && $# -ne 6 ]]; then
  echo ""
  echo "========================="
  echo "Usage: ./run.sh [WORLD_SIZE] [BATCH_SIZE] [EPOCH_SIZE] [LEARNING_RATE] [GPUSET] [DE(OPTIONAL)]"
  echo "========================="
  echo ""
  exit 0
fi
WORLD_SIZE=$1; BATCH_SIZE=$2; EPOCH_SIZE=$3; LEARNING_RATE=$4; GPUSET=$5; DE=${6:-false}
EXTRA=${DLB_EXTRA_ARGS:-}          # e.g. DLB_EXTRA_ARGS="--synthetic true --throttle_rank 3 --throttle_ms 5"
cd "$(dirname "$0")"
for dbs in true false; do
  for dataset in cifar10 cifar100; do
    for model in resnet densenet googlenet regnet; do
      CMD="python dbs.py -d false -ws $WORLD_SIZE -lr $LEARNING_RATE -b $BATCH_SIZE -e $EPOCH_SIZE -ds $dataset -dbs $dbs -m $model -ocp true -gpu $GPUSET -de $DE $EXTRA"
      echo ""; echo "========================="; echo "Running:"; echo "$CMD"; echo "========================="; echo ""
      if ! eval "$CMD"; then
        echo ""; echo "========================="; echo "FAILED AT DATASET $dataset, MODEL $model"; echo "========================="; echo ""
        exit 1
      fi
    done
  done
done
