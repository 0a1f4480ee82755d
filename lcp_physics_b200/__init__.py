"""lcp_physics_b200 -- B200-native batched LCP contact solver behind the
`lcp_physics` API (LCPFunction / PdipmEngine). See DESIGN.md."""
from .lcp import LCPFunction, solve_forward, solve_backward  # noqa: F401

__all__ = ["LCPFunction", "solve_forward", "solve_backward"]
