/*
 * Link-time stand-ins (drop-in CLI and test builds, SURVEY.md section 8b "build-time stubs") for the CPLEX-backed entry points
 * that the reference's src/main.c:41-76 dispatches to.  They let the UNCHANGED driver link
 * without CPLEX; every one reports UNIMPLEMENTED (errors.h:46), which main.c logs as fatal.
 * None of these is on the hot path.
 */
#include <stdbool.h>

enum { STUB_UNIMPLEMENTED = 12 };   /* ERROR_CODE UNIMPLEMENTED, src/utils/errors.h:46 */

int cx_Nosec(void) { return STUB_UNIMPLEMENTED; }
int cx_BendersLoop(bool patching) { (void)patching; return STUB_UNIMPLEMENTED; }
int cx_BranchAndCut(void) { return STUB_UNIMPLEMENTED; }
int mh_HardFixing(void) { return STUB_UNIMPLEMENTED; }
int mh_LocalBranching(void) { return STUB_UNIMPLEMENTED; }
