"""lcp_physics_b200 -- B200-native batched LCP contact solver behind the
`lcp_physics` API (LCPFunction / PdipmEngine). See DESIGN.md."""
from .lcp import LCPFunction, solve_forward, solve_backward  # noqa: F401

__all__ = ["LCPFunction", "solve_forward", "solve_backward"]
# fused engine path (contact list in, solution out): lcp_physics_b200.engines.engine_solve / B200PdipmEngine
