"""`python -m mpassit_b200 <namelist>` -- the reference's `mpassit <namelist>` (mpassit.F90:23-146) through the C++
mirror (host/run.cpp): one process per GPU; under `torchrun` the ranks share the output file and the communicator
stands in for MPI.  Without an argument the namelist is ./fort.41, as in the reference (mpassit.F90:53-66).
The var-list files (diaglist, histlist_2d, histlist_3d, histlist_soil) are read from the working directory."""
from __future__ import annotations

import os
import sys


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    namelist = argv[0] if argv else "fort.41"
    if not argv:
        print(" no namelist entry provided at execution, defaulting to using fort.41")
    if not os.path.exists(namelist):
        print(f" FATAL ERROR: namelist file - {namelist} does not exist.", file=sys.stderr)
        return 1
    from . import build, host

    build.build_all()
    host.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")
        comm = host.torch_comm()
    if rank == 0:
        print("- NPETS IS  ", world)
    try:
        st = host.run(namelist, None, device=local, rank=rank, nranks=world, comm=comm)
    except host.HostError as e:  # error_handler, utils.F90:16-33: message, then a non-zero exit
        print(f" FATAL ERROR: {e.msg}", file=sys.stderr)
        print(f" IOSTAT IS: {e.rc}", file=sys.stderr)
        return 999
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
    if rank == 0:
        print(f"- setup {st.setup_ms:.0f} ms, read {st.read_ms:.0f} ms, interp {st.interp_ms:.0f} ms, write {st.write_ms:.0f} ms")
    print("- DONE.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
