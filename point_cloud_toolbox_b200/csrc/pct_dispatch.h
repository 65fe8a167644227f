// Limits of the query kernels.
#pragma once

#define PCT_MAX_K 128

// extra neighbour-list slots for the boundary bin of the distance histogram (ties with the k-th neighbour)
#define PCT_TIE_SLACK 16
