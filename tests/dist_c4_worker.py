"""Worker of test_c4_sharded_scene_nccl (launched by torch.distributed.run, one rank per GPU): BASELINE config[3] /
SURVEY C4 -- one scene block-partitioned over the ranks, ONE NCCL all-gather of the packed codes, match on the gathered
table on every rank; the gathered table and the assignments must equal a 1-GPU run of the full list bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import livingscenes_b200 as ls
    from conftest import SHIPPED_WEIGHTS
    from livingscenes_b200 import synthetic as S
    from livingscenes_b200.dist import encode_sharded

    n_scene = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    sd = torch.load(SHIPPED_WEIGHTS, map_location="cpu", weights_only=True) if os.path.exists(SHIPPED_WEIGHTS) \
        else S.random_state_dict(0)
    model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
    half = n_scene // 2
    ref = S.synth_parts(half, 1024, 77)
    g = torch.Generator().manual_seed(78)
    perm = torch.randperm(half, generator=g)
    x = torch.cat([ref, S.random_rotations(half, 79) @ ref[perm] + torch.randn(half, 3, 1, generator=g)]).to(dev)
    code = encode_sharded(model, x)                       # sharded encode + all-gather
    full = model.encode(x)                                # the same list on ONE GPU
    for k in ("z_so3", "z_inv", "s", "t"):
        assert torch.equal(code[k].reshape(-1), full[k].reshape(-1)), f"rank {rank}: gathered {k} differs from the 1-GPU run"
    m = ls.sequential_matcher(code["z_inv"][:half].contiguous(), code["z_inv"][half:].contiguous())
    m1 = ls.sequential_matcher(full["z_inv"][:half].contiguous(), full["z_inv"][half:].contiguous())
    assert torch.equal(m["matches0"], m1["matches0"]) and torch.equal(m["matches1"], m1["matches1"])
    assert torch.equal(m["matches0"].cpu(), torch.argsort(perm)), "planted permutation not recovered"
    # every rank must hold the SAME table
    ref_tab = code["z_inv"].clone()
    dist.broadcast(ref_tab, 0)
    assert torch.equal(ref_tab, code["z_inv"])
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"C4 OK world={world} n={n_scene}")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
