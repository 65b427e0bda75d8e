/* harness_patching.c -- TEST driver (built by csrc/Makefile.interpose into build/bin/) for cx_patching
 * (cplex_model.c:693-755), which the reference CLI only reaches through CPLEX (cx_BendersLoop, :205-212).
 * Linked exactly like the drop-in CLI (interposer + libtspcuda ahead of the unmodified reference objects):
 * the reference-side call below goes through the PLT and lands in the interposer's cx_patching.
 *   usage: harness_patching N SEED solution.i32   (file: N successors then N component labels, int32)
 *   prints: <patches> <cost as C99 hex float> <FNV-1a-64 of path then comp>
 * It repeats the caller's loop `while (solution.ncomp > 1) cx_patching(&solution)`. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

int refapi_random_points(int n, int seed);                            /* oracle/ref_api.c: srand + tsp_generate_randompoints */
int refapi_patching(int *path, int *comp, int *ncomp, double *cost);  /* oracle/ref_api.c: cx_patching(&solution) */
double refapi_solution_cost(const int *path);

int main(int argc, char **argv) {
    if (argc != 4) return 2;
    int n = atoi(argv[1]), seed = atoi(argv[2]);
    int *buf = (int *)malloc(sizeof(int) * 2 * (size_t)n);
    FILE *f = fopen(argv[3], "rb");
    if (!buf || !f || fread(buf, sizeof(int), 2 * (size_t)n, f) != 2 * (size_t)n) return 3;
    fclose(f);
    if (refapi_random_points(n, seed) != 0) return 4;
    int *path = buf, *comp = buf + n;
    int ncomp = 0;
    for (int i = 0; i < n; i++) if (comp[i] > ncomp) ncomp = comp[i];     /* labels 1..ncomp */
    double cost = refapi_solution_cost(path);
    int patches = 0;
    while (ncomp > 1) { refapi_patching(path, comp, &ncomp, &cost); patches++; }
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < 2 * n; i++) { h ^= (uint64_t)(uint32_t)buf[i]; h *= 0x100000001b3ULL; }
    printf("%d %a %016llx\n", patches, cost, (unsigned long long)h);
    fflush(stdout);
    return 0;
}
