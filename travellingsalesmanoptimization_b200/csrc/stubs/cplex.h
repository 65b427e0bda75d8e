/*
 * Build-time stand-in for IBM CPLEX's <cplex.h> (needed wherever the reference driver is compiled without CPLEX: the drop-in CLI and the test builds).
 *
 * The reference's driver (src/main.c) includes algorithms/matheuristics.h:1 -> <cplex.h>
 * only to see a handful of opaque handle types in prototypes (cplex_model.h:16-18,44-88).
 * CPLEX itself is proprietary and absent from this image; the exact / matheuristic code is
 * out of scope (north_star: "may be stubbed at build time when the library is absent").
 * Only the opaque types the headers mention are declared here; no CPLEX entry point exists.
 */
#ifndef TSPCUDA_CPLEX_STUB_H_
#define TSPCUDA_CPLEX_STUB_H_

typedef struct tspcuda_stub_cpxenv *CPXENVptr;
typedef struct tspcuda_stub_cpxlp *CPXLPptr;
typedef struct tspcuda_stub_cpxcb *CPXCALLBACKCONTEXTptr;
typedef long long CPXLONG;
#define CPXPUBLIC

#endif
