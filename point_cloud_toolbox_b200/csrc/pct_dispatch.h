// Limits of the query kernels.
#pragma once

#define PCT_MAX_K 128

// extra neighbour-list slots for the boundary bin of the distance histogram (ties with the k-th neighbour)
#define PCT_TIE_SLACK 16

// staged kNN kernel: queries (threads) per CTA, and the resident CTAs per SM it is compiled for
// (register cap) and sized for (shared memory)
#ifndef PCT_STAGED_BLOCK
#define PCT_STAGED_BLOCK 256
#endif
#ifndef PCT_STAGED_CTAS
#define PCT_STAGED_CTAS 3
#endif

// candidates per trip of the staged candidate loop
#ifndef PCT_SCAN_WIDTH
#define PCT_SCAN_WIDTH 4
#endif

// pass 1 of the selection without a branch around the histogram update
#ifndef PCT_BRANCHFREE_HIST
#define PCT_BRANCHFREE_HIST 1
#endif

// pass 2 of the selection with one predicated store instead of nested branches
#ifndef PCT_BRANCHFREE_PART
#define PCT_BRANCHFREE_PART 1
#endif

// pass 2 of the selection skips the cells of the staged block that lie beyond the boundary bin.  Off: 41 % of the
// pass-2 candidates go away on the benchmark surface, but the warp pays whole 4-wide trips of its slowest lane and
// the per-row bound tests of every lane, and the kernel got 9 % slower (12.5 vs 11.4 ms at 20 M points, k = 20)
#ifndef PCT_CULL_PASS2
#define PCT_CULL_PASS2 0
#endif

// pass 1 updates its byte histogram with one shared-memory reduction per candidate instead of a byte
// load-add-store (not measured yet: profiles/README.md, end of r01u).  Only for kernels whose scratch is shared memory.
#ifndef PCT_HIST_RED
#define PCT_HIST_RED 0
#endif

// bins of the distance histogram of pass 1 (a multiple of 4): fewer bins = fewer words to clear and to prefix-sum
// per query, a wider boundary bin = more candidates for the fp64 re-rank
#ifndef PCT_HIST_BINS
#define PCT_HIST_BINS 64
#endif

// one-pass selection in the staged kernel (knn_select<.., ONEPASS>): pass 1 lists the candidates below a cut estimated
// from the local density and the histogram is built from that list; queries whose cut missed go to the L1/L2 kernel.
// Not measured yet; the build then sets IndexView::cut_gain to 6.3 = 2.2 * 9 / pi.
#ifndef PCT_ONEPASS
#define PCT_ONEPASS 0
#endif
