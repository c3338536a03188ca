/* tpp_foamrun.c - `foamRun` for a case directory from a plain C host: no Python, no FoamFile
 * parser, nothing but dlopen and five entry points of include/tppvof.h.  It is what
 *     make run / make resume  ->  foamRun        (/root/reference/circularSloshingTank/Makefile:85,98)
 * needs from the library, and shows that a non-Python orchestrator (the reference's main.py:333-348
 * only needs an executable that exits 0 or not) can bind the boundary as it stands.
 *
 *   gcc -Iinclude tools/tpp_foamrun.c -o tpp_foamrun -ldl
 *   ./tpp_foamrun openfoam-tpp_b200/libtppvof.so -case <dir> [-device N] [-maxSteps N] [-interface]
 *
 * The library is named on the command line (not linked) so that the tests can hand it the host
 * emulation build; a deployment links libtppvof.so directly.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tppvof.h"

int main(int argc, char** argv) {
    const char* lib = argc > 1 ? argv[1] : NULL;
    const char* dir = ".";
    int device = 0;
    long max_steps = -1;
    int flags = TPP_RUN_LOG;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-case") && i + 1 < argc) dir = argv[++i];
        else if (!strcmp(argv[i], "-device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-maxSteps") && i + 1 < argc) max_steps = atol(argv[++i]);
        else if (!strcmp(argv[i], "-interface")) flags |= TPP_RUN_INTERFACE;
        else if (!strcmp(argv[i], "-noFunctionObjects")) continue;
        else {
            fprintf(stderr, "tpp_foamrun: unknown option %s\n", argv[i]);
            return 2;
        }
    }
    if (!lib) {
        fprintf(stderr, "usage: tpp_foamrun <libtppvof.so> [-case dir] [-device N] [-maxSteps N]\n");
        return 2;
    }
    void* h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "tpp_foamrun: %s\n", dlerror());
        return 2;
    }
    int (*open_)(const char*, int, int, tpp_handle*) = (int (*)(const char*, int, int, tpp_handle*))dlsym(h, "tpp_open");
    long (*run_)(tpp_handle, long, int) = (long (*)(tpp_handle, long, int))dlsym(h, "tpp_run_case");
    long (*query_)(tpp_handle, const char*, char*, long) = (long (*)(tpp_handle, const char*, char*, long))dlsym(h, "tpp_case_query");
    int (*destroy_)(tpp_handle) = (int (*)(tpp_handle))dlsym(h, "tpp_destroy");
    const char* (*error_)(void) = (const char* (*)(void))dlsym(h, "tpp_last_error");
    if (!open_ || !run_ || !query_ || !destroy_ || !error_) {
        fprintf(stderr, "tpp_foamrun: %s does not export the case entry points of tppvof.h\n", lib);
        return 2;
    }
    tpp_handle s = NULL;
    if (open_(dir, -1, device, &s) != 0) {
        fprintf(stderr, "--> FOAM FATAL ERROR: %s\n", error_());
        return 1;
    }
    char start[256];
    query_(s, "start_time", start, sizeof start);
    printf("Starting time loop from %s (%ld cells)\n", start, query_(s, "n_cells", NULL, 0));
    long steps = run_(s, max_steps, flags);
    if (steps < 0) {
        fprintf(stderr, "--> FOAM FATAL ERROR: %s\n", error_());
        destroy_(s);
        return 1;
    }
    destroy_(s);
    printf("End  (%ld steps)\n", steps);
    return 0;
}
