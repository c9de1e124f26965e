"""How fast can the GPUs of one box drain results into host memory AT THE SAME TIME?  (torchrun, one process per GPU)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_aggregate.py

Every rank copies its share of a 276 MB result (the images of a 10 000-frame SXI sweep) device -> host, all ranks at
once, (a) into its own cudaHostAlloc'ed buffer, (b) into its interleaved frames of ONE shared-memory buffer that every
rank page-locked (what bench.py's end-to-end arm does at N > 1).  Prints the aggregate GB/s of both: the ceiling of any
end-to-end number on this box, whatever the kernels do.
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist

    from birp_cpp_b200 import capi, sharding
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    capi.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nframes, H, W = 10000, 64, 108
    mine = nframes // world
    d = torch.zeros((mine, H, W), dtype=torch.float32, device=dev)
    own = torch.empty((mine, H, W), dtype=torch.float32, pin_memory=True)
    shared = sharding.SharedHostResult((nframes, H, W), rank, world)
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    fb = H * W * 4
    stream = torch.cuda.current_stream().cuda_stream

    def copy_own():
        own.copy_(d, non_blocking=True)

    def copy_shared():
        # what the library's strided sweep does: frame j of this rank -> slot rank + j * world of the shared buffer
        rc = rt.cudaMemcpy2DAsync(shared.array.ctypes.data + rank * fb, world * fb, d.data_ptr(), fb, fb, mine, 2, stream)
        assert rc == 0, rc

    out = {}
    for name, fn in (("own_pinned_buffers", copy_own), ("one_shared_buffer_interleaved", copy_shared)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        secs = time.perf_counter() - t0
        t = torch.tensor([secs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_GBs"] = round(reps * nframes * H * W * 4 / float(t[0]) / 1e9, 1)
    if rank == 0:
        out.update(n_gpus=world, bytes_per_pass=nframes * H * W * 4,
                   frames_per_s_ceiling=round(out["one_shared_buffer_interleaved_GBs"] * 1e9 / (H * W * 4)))
        print(json.dumps(out))
    shared.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
