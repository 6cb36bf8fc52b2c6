"""torchrun worker for tests/test_gpu_multi.py: partitioned operator / PCG on N ranks against the same problem
solved serially on this rank's own GPU (b200pa.selfcheck.partitioned_vs_serial - the same check bench.py runs
at world > 1 over the communicator of its timed region)."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "cardiac-ablation-ecm2_b200"))
import b200pa  # noqa: E402
from b200pa import selfcheck  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # B200PA_TEST_ONE_GPU=1: every rank on cuda:0 (a 1-GPU box still exercises the peer-memory kernels: the ranks'
    # mailboxes are mapped with CUDA IPC exactly as between GPUs, the GPU time-slices the processes); NCCL refuses
    # ranks that share a device, so the rendezvous is gloo and the communicator is created without NCCL
    one_gpu = os.environ.get("B200PA_TEST_ONE_GPU", "0") == "1"
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ctx = b200pa.Context(local)
    ids = [b200pa.Comm.unique_id() if (rank == 0 and not one_gpu) else None]
    dist.broadcast_object_list(ids, src=0)
    comm = b200pa.Comm(ctx, ids[0], rank, world)
    r = selfcheck.partitioned_vs_serial(ctx, comm, rank, world, p)
    want_p2p = os.environ.get("B200PA_NO_P2P", "0") != "1"
    assert comm.p2p_enabled() == want_p2p, f"transport: peer-memory path enabled={comm.p2p_enabled()}, expected {want_p2p}"
    g = selfcheck.multigrid_partitioned_vs_serial(ctx, comm, rank, world, orders=(1, 2, 3) if p != 3 else (1, 3))
    if rank == 0:
        print(f"MULTI_OK world={world} p={p} p2p={comm.p2p_enabled()} apply={r['apply_rel_err']:.2e} diag={r['diag_rel_err']:.2e} "
              f"pcg={r['pcg15_rel_err']:.2e} cheb={r['cheb_pcg6_rel_err']:.2e} fact={r['factorised_apply_rel_err']:.2e} "
              f"its={r['pcg_iters_to_1e-8']} mg_transfer={g['transfer_rel_err']:.2e} mg_vcycle={g['vcycle_rel_err']:.2e} "
              f"mg_pcg={g['mgpcg_rel_err']:.2e} mg_its={g['mgpcg_iters']}", flush=True)
    comm.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
