"""Launched by torchrun on N >= 2 GPUs (tests/test_parity_gpu.py::test_nccl_ranks_equal_single_gpu):
the N-rank NCCL run (k_pack -> ncclSend/Recv over NVLink -> k_unpack) must leave OWNED entities
bit-identical to a single-partition run of the same library on one GPU, and within 1e-12 of the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpas_regent_b200 import _abi, dynamics, icosa, init_jw, parallel  # noqa: E402
from tests.test_parallel import _assert_owned_equal  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1] if len(sys.argv) > 1 else "literal"
    # modes: literal | overlap | corrected  = the host-side (Python) schedule;  native | native_corrected | native_scalars = the
    # schedule inside the library (mpasb200_srk3_dist)
    native = mode.startswith("native")
    corrected, overlap = mode.endswith("corrected"), mode == "overlap"
    scalars = mode == "native_scalars"
    L, dt, steps = 12, 400.0, (1 if corrected else 3)     # corrected physics: the first step is the finite one
    mesh = icosa.make_icosahedral_mesh(2562)
    st = init_jw.make_state(mesh, L, _abi.INDEX_CORRECTED)
    cfg = _abi.default_config(rkarg_policy=_abi.RKARG_STAGE_INDEX, device=local,
                              physics_mode=_abi.PHYSICS_CORRECTED if corrected else _abi.PHYSICS_LITERAL,
                              config_scalar_advection=int(scalars))
    if scalars:
        rng = np.random.default_rng(17)
        st.f["scalars"] = 1e-3 * (1.0 + rng.random((mesh.nCells, L + 1, 8)))
    stream = torch.cuda.Stream()
    sh = parallel.make_shards(st, world)[rank]
    lm = sh["lm"]
    d = dynamics.Dynamics(_abi.make_dims(len(lm.cells), len(lm.edges), len(lm.vertices), L), cfg)
    d.set_stream(stream.cuda_stream)
    d.upload_mesh(sh["static"]); d.upload_state(sh["f"], sh["vert"])
    run = (parallel.NativeDistributedDynamics(d, lm, rank, world) if native
           else parallel.DistributedDynamics(d, parallel.NcclExchanger(d, lm, stream)))
    if native:
        c0, c1 = d.class_range(_abi.CELL, 0), d.class_range(_abi.CELL, 1)
        assert c1[1] - c1[0] > 0 and c0[1] - c0[0] > 0
    if overlap:       # acoustic-loop exchanges on the communication stream, interior compute underneath
        assert run.enable_overlap()
        c0, c1 = d.class_range(_abi.CELL, 0), d.class_range(_abi.CELL, 1)
        assert c1[1] - c1[0] > 0 and c0[1] - c0[0] > 0
    run.init_diagnostics()
    for _ in range(steps):
        run.step(dt)
    run.flush()
    torch.cuda.synchronize()
    single = dynamics.Dynamics(dynamics.dims_of(mesh, L), cfg)
    single.upload_mesh(st.static); single.upload_state(st.f, st.vert)
    single.atm_compute_solve_diagnostics(False, -1)
    for _ in range(steps):
        single.atm_srk3(dt)
    _assert_owned_equal(single, d, lm)
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK mode={mode} world={world} owned_cells={lm.n_owned[0]} ghosts={len(lm.cells) - lm.n_owned[0]}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
