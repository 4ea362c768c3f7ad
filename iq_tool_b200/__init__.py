"""iq_tool_b200 — B200-native (sm_100a) implementation of pclov3r/iq_tool's per-block
sample-processing chain behind the reference's stage/module API.  See DESIGN.md."""
from .configs import ChainConfig, Workload, baseline_workloads  # noqa: F401

__all__ = ["ChainConfig", "Workload", "baseline_workloads"]
