"""TEST INFRASTRUCTURE ONLY (never imported by npore_b200/).

`import oracle` finds this package when the repository root precedes oracle/ on sys.path and oracle/oracle.py otherwise
(tests/conftest.py puts both there; spawned worker processes inherit whatever order the parent ended up with).  Both must
answer alike, so the package re-exports the module."""
from .oracle import *  # noqa: F401,F403
