"""Does this box expose NVLink multicast (NVLS) through torch's symmetric memory? (2+ GPUs, torchrun)"""
import os

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
try:
    import torch.distributed._symmetric_memory as symm_mem

    t = symm_mem.empty(1 << 20, dtype=torch.float32, device="cuda")
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
    info = {"rank": rank, "world": hdl.world_size, "buffer_ptrs": len(hdl.buffer_ptrs), "multicast_ptr": hex(getattr(hdl, "multicast_ptr", 0) or 0),
            "signal_pads": len(hdl.signal_pad_ptrs)}
    print("SYMM", info, flush=True)
    # functional check of the in-switch reduction through torch's own op, if present
    t.fill_(float(rank + 1))
    hdl.barrier()
    if getattr(hdl, "multicast_ptr", 0):
        try:
            torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
            torch.cuda.synchronize()
            print("SYMM multimem_all_reduce_ ->", float(t[0]), flush=True)
        except Exception as e:  # noqa: BLE001
            print("SYMM multimem op failed:", repr(e)[:200], flush=True)
except Exception as e:  # noqa: BLE001
    print("SYMM unavailable:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
