// Limits of the query kernels.
#pragma once

#define PCT_MAX_K 128

// two-pass selection: extra neighbour-list slots for the boundary bin of the distance histogram (ties with the k-th neighbour)
#define PCT_TIE_SLACK 16

// staged kernels: queries (threads) per CTA, and the resident CTAs per SM they are compiled for
// (register cap) and sized for (shared memory).  384 x 2 measured 2 % slower at k = 20 and 3 % faster at k = 32
// (profiles/variants_r02a.txt); 256 x 3 is kept.
#ifndef PCT_STAGED_BLOCK
#define PCT_STAGED_BLOCK 256
#endif
#ifndef PCT_STAGED_CTAS
#define PCT_STAGED_CTAS 3
#endif

// staging copy of the staged kernels: 1 = TMA bulk copies (cp.async.bulk global -> shared, one per cell run, completion
// on an mbarrier), 0 = 16-byte loads through registers
#ifndef PCT_TMA_STAGE
#define PCT_TMA_STAGE 1
#endif
