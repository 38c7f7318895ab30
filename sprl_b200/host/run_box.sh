#!/bin/bash
# run_box.sh -- one worker process per GPU of this box, the way the reference's cluster script starts one per CPU core.
#   sprl_b200/host/run_box.sh <OTHWorker|C4Worker|GoWorker> <first_task_id> <num_tasks> [gpus]
# GPU g runs `<worker> <first_task_id + g> <num_tasks>` with SPRL_DEVICE=g, so every GPU writes its own
# data/games/<run>/<group>/<task>/ directory exactly as a reference worker task would (the controller's path logic,
# scripts/othello_controller.py:78-80, needs no change).  Games never share a stream: task t plays stream ids
# t, t + num_tasks, ...  Search and batch sizes come from the SPRL_* environment variables of worker_main.hpp.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
WORKER=${1:?worker name}; FIRST=${2:?first task id}; TASKS=${3:?number of tasks}
GPUS=${4:-$(nvidia-smi -L | wc -l)}
pids=()
for ((g = 0; g < GPUS; ++g)); do
  SPRL_DEVICE=$g "$HERE/bin/$WORKER" $((FIRST + g)) "$TASKS" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=$?; done
exit $rc
