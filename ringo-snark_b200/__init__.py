"""ringo-snark_b200 -- B200 (sm_100a) implementation of Ringo-SNARK's polynomial-arithmetic
hot path, behind the C ABI of include/ringo_cuda.h.

This Python package is only the ctypes binding of that ABI (the same entry points a Go
maintainer binds with cgo) plus thin classes that mirror the reference's
operator interface (`bigpoly.CyclicEvaluator`, `bigpoly.CyclotomicEvaluator`, `bigpoly.Poly`)
so the parity tests read like the reference's own.  There is no CPU fallback: importing
works anywhere, but every compute call needs libringo_b200.so and a B200.
"""
from . import capi            # noqa: F401
from .capi import RingoError, lib, load, check  # noqa: F401
from . import jindo, jindo_host, parallel, buckler, circuits  # noqa: F401
from .bigpoly import (Field, Batch, Poly, CyclicTransformer, CyclotomicTransformer,  # noqa: F401
                      CyclicEvaluator, CyclotomicEvaluator, MODULI, modulus)
