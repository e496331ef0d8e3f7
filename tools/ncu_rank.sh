#!/bin/bash
# One rank of a torchrun job under ncu, SINGLE-PASS metrics only (duration + NVLink bytes): a kernel that exchanges data with its
# peer cannot be replayed (the peer does not send twice), so no --set full here.  Launched as
#   python -m torch.distributed.run --nproc-per-node 2 --no-python bash tools/ncu_rank.sh <out prefix> <python script> [args]
out=$1; shift
exec ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum --clock-control none --replay-mode kernel \
    -k regex:'p2p_halo_ll_kernel|pcg_update_p2p_kernel|pa_apply_eo_kernel' -c 400 --csv --log-file ${out}_rank${LOCAL_RANK}.csv python "$@"
