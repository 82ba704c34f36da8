// sum_order.h -- the ONE place that fixes the association order of the three-term sums the reference evaluates through Eigen
// expression templates (dot products, squared norms, matrix-vector and matrix-matrix products).  Eigen is not on this image,
// so its order cannot be inspected: FCL_SUM3_ORDER = 0 (default) is left to right, (x0 + x1) + x2; 1 is
// x0 + (x1 + x2).  The oracle honours the same macro (oracle/fcl_oracle_vec.hpp `sum3`), so the day a real libfcl can be run
// next to this library the order is flipped on both sides with one flag and re-verified: tests/test_sum_order_hook.py builds
// both sides with -DFCL_SUM3_ORDER=1 and re-runs the bit-parity tests.  Sums the reference writes out in scalar C++
// (covariance accumulation, the SAT's four-term radii, ...) have the order of the source text and do not go through here.
#pragma once
#ifndef FCL_SUM3_ORDER
#define FCL_SUM3_ORDER 0
#endif
// A macro, not a function: with the default order the expansion is token for token the expression that was written out
// before, so the compiled kernels do not change by a single instruction (checked on the SASS of the whole library).
#if FCL_SUM3_ORDER == 0
#define FCL_SUM3(x0, x1, x2) (((x0) + (x1)) + (x2))
#else
#define FCL_SUM3(x0, x1, x2) ((x0) + ((x1) + (x2)))
#endif
