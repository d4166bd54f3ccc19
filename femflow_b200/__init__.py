"""femflow_b200 -- B200-native (sm_100a) MLS/APIC MPM substep behind FEMFlow's
solver / simulation API.  Importing the package does not need a GPU; creating a
solver does, and fails loudly without the compiled CUDA library."""

__version__ = "0.1.0"

from ._build import build_library  # noqa: F401
