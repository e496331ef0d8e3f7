#!/bin/bash
# One rank of a torchrun job under ncu with SINGLE-PASS metrics only: a kernel that exchanges data with its peer cannot be
# replayed (the peer does not send twice), so no --set full here.  Launched as
#   python -m torch.distributed.run --nproc-per-node 2 --no-python bash tools/ncu_rank.sh <mode> <out prefix> <python script> [args]
# mode "time": every rank under ncu, gpu__time_duration.sum of the exchange-carrying kernels;
# mode "nvl":  rank 0 alone under ncu, NVLink byte counters next to the duration (two ncu instances reading the NVLink
#              counters at once failed with UnknownError on the 2-GPU box).
mode=$1; out=$2; shift 2
K='regex:p2p_halo_ll_kernel|pcg_update_p2p_kernel|pa_apply_eo_kernel'
if [ "$mode" = time ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file ${out}_rank${LOCAL_RANK}.csv python "$@"
elif [ "$LOCAL_RANK" = 0 ]; then
  exec ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum --clock-control none -k "$K" -c 600 --csv --log-file ${out}_rank0.csv python "$@"
else
  exec python "$@"
fi
