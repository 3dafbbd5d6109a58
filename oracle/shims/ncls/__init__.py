"""shim of ``ncls`` over the CPU oracle (test infrastructure)."""
from oracle.oracle import NCLS  # noqa: F401
