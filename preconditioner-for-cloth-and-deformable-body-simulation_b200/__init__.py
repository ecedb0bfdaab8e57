"""B200-native multilevel additive Schwarz (MAS) preconditioner: drop-in for the one hot path of
V-Sekai/preconditioner-for-cloth-and-deformable-body-simulation (class SE::SeSchwarzPreconditioner).

  csrc/           hand-written sm_100a CUDA kernels + the extern "C" boundary (include/mas_b200.h); the batched inversion's
                  device code (mas_invert.cuh) is also compiled for the host by tests/emu (thread-block emulation)
  schwarz.py      host-side mirror of the reference class over that C ABI (ctypes)
  synth.py        deterministic synthetic inputs for the BASELINE.json configs
  collide.py      proximity-based EF / EE / VF stencil producer (caller-side workload generation, torch tensor ops)
  pcg.py          caller-side PCG loop (the reference ships none) over mas_pcg_solve, used for iteration-count parity
  partition.py    multi-GPU host side: Morton-contiguous shards, begin -> exchange -> end driver

The directory name contains hyphens; import it with
    importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
"""
from .schwarz import SeSchwarzPreconditioner, MasError, load_library, LIB_PATH, EXPORTS  # noqa: F401
from . import synth  # noqa: F401
from . import partition  # noqa: F401
from . import collide  # noqa: F401
from .pcg import pcg_solve, PcgResult  # noqa: F401
