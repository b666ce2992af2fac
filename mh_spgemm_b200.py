"""Import shim: the package directory is ``mh-spgemm_b200/`` (not a valid Python
identifier), so this module exposes it as ``mh_spgemm_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mh-spgemm_b200")]
from mh_spgemm_b200.csr import CSR  # noqa: E402,F401
