/* A host harness in plain C99 on top of the C ABI only (include/qaintensor_cuda.h): circuit file -> tensor network
 * (qtn_net_create + qtn_net_tensor_circuit, the reference's tensor_circuit!, src/tensor_circuit.jl:44-51) -> closed
 * amplitude network -> contraction order (reference treewidth heuristic or the searched extension) -> contract on the GPU.
 *
 *   gcc -std=c99 -I include examples/amplitude.c -L qaintensor.jl_b200/lib -lqaintensor_cuda -Wl,-rpath,... -o amplitude
 *   ./amplitude circuit.bin [treewidth|search] [max_log2_elems]
 *
 * circuit.bin (little endian): int32 nq, int32 ngates; per gate: int32 wire1, wire2 (1-based) and 16 complex doubles
 * (4x4, column-major); then nq int32 output bits.  Exit codes: 0 ok, 2 bad input, 3 library error (message printed). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qaintensor_cuda.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != 0) {                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, qtn_last_error()); \
            return 3;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s circuit.bin [treewidth|search] [max_log2_elems]\n", argv[0]); return 2; }
    const int method = (argc > 2 && strcmp(argv[2], "search") == 0) ? 1 : 0;
    const int max_log2 = argc > 3 ? atoi(argv[3]) : -1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t nq = 0, ng = 0;
    if (fread(&nq, 4, 1, f) != 1 || fread(&ng, 4, 1, f) != 1 || nq < 1 || nq > 60 || ng < 0) { fprintf(stderr, "bad header\n"); return 2; }
    int32_t* wires = malloc(sizeof(int32_t) * 2 * (size_t)(ng > 0 ? ng : 1));
    int32_t* nwires = malloc(sizeof(int32_t) * (size_t)(ng > 0 ? ng : 1));
    double* mats = malloc(sizeof(double) * 32 * (size_t)(ng > 0 ? ng : 1));
    const void** mptr = malloc(sizeof(void*) * (size_t)(ng > 0 ? ng : 1));
    int32_t* bits = malloc(sizeof(int32_t) * (size_t)nq);
    for (int g = 0; g < ng; ++g) {
        if (fread(wires + 2 * g, 4, 2, f) != 2 || fread(mats + 32 * g, 8, 32, f) != 32) { fprintf(stderr, "truncated gate %d\n", g); return 2; }
        nwires[g] = 2;
        mptr[g] = mats + 32 * g;
    }
    if (fread(bits, 4, (size_t)nq, f) != (size_t)nq) { fprintf(stderr, "truncated bits\n"); return 2; }
    fclose(f);

    /* |0>^nq: nq rank-1 tensors, every leg open */
    const double ket0[4] = {1.0, 0.0, 0.0, 0.0};
    const int64_t two = 2;
    const void** kdata = malloc(sizeof(void*) * (size_t)nq);
    int32_t* ranks = malloc(sizeof(int32_t) * (size_t)nq);
    const int64_t** dims = malloc(sizeof(int64_t*) * (size_t)nq);
    int32_t* openidx = malloc(sizeof(int32_t) * 2 * (size_t)nq);
    for (int i = 0; i < nq; ++i) { kdata[i] = ket0; ranks[i] = 1; dims[i] = &two; openidx[2 * i] = i + 1; openidx[2 * i + 1] = 1; }
    qtn_net* net = NULL;
    CHECK(qtn_net_create(nq, kdata, ranks, dims, 0, NULL, nq, openidx, &net));
    CHECK(qtn_net_tensor_circuit(net, ng, nwires, wires, mptr));
    CHECK(qtn_net_close(net, bits));
    int32_t sizes[3];
    CHECK(qtn_net_sizes(net, sizes));
    printf("network: %d tensors, %d contractions, %d open legs\n", sizes[0], sizes[1], sizes[2]);
    CHECK(qtn_net_optimize_order(net, method, 128, 0, max_log2));
    printf("order ok (%s)\n", method ? "search" : "treewidth");
    fflush(stdout);
    double amp[2] = {0.0, 0.0};
    int32_t out_rank = -1;
    int64_t out_dims[64];
    CHECK(qtn_net_contract(net, QTN_C128, max_log2, amp, &out_rank, out_dims));
    printf("amplitude %.17g %.17g\n", amp[0], amp[1]);
    CHECK(qtn_net_destroy(net));
    return 0;
}
