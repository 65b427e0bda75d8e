/*
 * TEST INFRASTRUCTURE (oracle/ only): a second stand-in for <cplex.h>, used for ONE purpose -- to let the
 * reference's src/algorithms/cplex_model.c compile unchanged so that its plain-C helper cx_patching
 * (cplex_model.c:693-755) can be called as the parity reference of tspcuda_patching.
 * cx_patching touches no CPLEX entry point.  The macro values below are arbitrary distinct integers:
 * nothing that reads them is ever executed (no CPX* function exists in this image; oracle/Makefile
 * turns every one of them into a stub that aborts).
 */
#ifndef TSPCUDA_ORACLE_CPLEX_CX_STUB_H_
#define TSPCUDA_ORACLE_CPLEX_CX_STUB_H_
#include "../cplex.h"

#define CPX_PARAM_TILIM 1
#define CPX_PARAM_SCRIND 2
#define CPX_PARAM_CLONELOG 3
#define CPX_OFF 0
#define CPX_ON 1
#define CPXMIP_TIME_LIM_FEAS 101
#define CPXMIP_TIME_LIM_INFEAS 102
#define CPXMIP_INFEASIBLE 103
#define CPXMIP_ABORT_FEAS 104
#define CPXMIP_ABORT_INFEAS 105
#define CPXMIP_OPTIMAL_TOL 106
#define CPXMIP_OPTIMAL 107
#define CPX_CALLBACKCONTEXT_CANDIDATE 0x20
#define CPX_CALLBACKCONTEXT_RELAXATION 0x40
#define CPX_INFBOUND 1.0E+20
#define CPX_MIPSTART_NOCHECK 4
#define CPXCALLBACKINFO_THREADID 0
#define CPXCALLBACKINFO_NODECOUNT 1
#define CPXCALLBACKINFO_NODEDEPTH 2
#define CPX_USECUT_FORCE 0
#define CPX_USECUT_PURGE 1
#define CPXCALLBACKSOLUTION_NOCHECK 5

#endif
