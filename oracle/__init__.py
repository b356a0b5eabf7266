"""CPU oracle (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; lexls_b200/ never does.  See oracle/lexls_oracle.hpp for the contract.
"""
from .pyoracle import (  # noqa: F401
    OracleParams,
    build,
    lse_solve_batch,
    lsi_solve_batch,
    lsi_solve_detail,
    set_regularization,
    set_least_norm,
)
