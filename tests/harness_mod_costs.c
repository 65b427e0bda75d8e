/* harness_mod_costs.c -- TEST driver (built by csrc/Makefile.interpose into build/bin/) for the one hot-path entry point the reference CLI cannot reach
 * without CPLEX: h_Greedy_2opt_mod_costs (heuristics.c:118-149, called from the relaxation callback,
 * cplex_model.c:1217).  Linked exactly like the drop-in CLI (interposer + libtspcuda ahead of the
 * unmodified reference objects), it sets nnodes through the reference's own generator, reads a
 * row-major matrix of doubles from a file and calls the reference symbol.
 *   usage: harness_mod_costs N matrix.f64        prints: <error code> <cost as C99 hex float> <FNV-1a-64 of the path> */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

int refapi_random_points(int n, int seed);                       /* oracle/ref_api.c: srand + tsp_generate_randompoints */
int refapi_mod_costs(double *costs, int *path, double *cost);    /* oracle/ref_api.c: h_Greedy_2opt_mod_costs */

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    int n = atoi(argv[1]);
    double *m = (double *)malloc(sizeof(double) * (size_t)n * n);
    int *path = (int *)malloc(sizeof(int) * n);
    FILE *f = fopen(argv[2], "rb");
    if (!m || !path || !f || fread(m, sizeof(double), (size_t)n * n, f) != (size_t)n * n) return 3;
    fclose(f);
    if (refapi_random_points(n, 1) != 0) return 4;
    double cost = 0;
    int e = refapi_mod_costs(m, path, &cost);
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < n; i++) { h ^= (uint64_t)(uint32_t)path[i]; h *= 0x100000001b3ULL; }
    printf("%d %a %016llx\n", e, cost, (unsigned long long)h);
    fflush(stdout);
    return 0;
}
