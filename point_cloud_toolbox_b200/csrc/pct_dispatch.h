// Compile-time neighbour-count buckets: the selection list lives in registers,
// so k is rounded up to the next bucket (k = 20 and k = 32 are exact).
#pragma once

#define PCT_MAX_K 128

#define PCT_DISPATCH_KT(k, ...)                                  \
    do {                                                         \
        if ((k) <= 8) { constexpr int KT = 8; __VA_ARGS__; }     \
        else if ((k) <= 16) { constexpr int KT = 16; __VA_ARGS__; } \
        else if ((k) <= 20) { constexpr int KT = 20; __VA_ARGS__; } \
        else if ((k) <= 24) { constexpr int KT = 24; __VA_ARGS__; } \
        else if ((k) <= 32) { constexpr int KT = 32; __VA_ARGS__; } \
        else if ((k) <= 50) { constexpr int KT = 50; __VA_ARGS__; } \
        else if ((k) <= 64) { constexpr int KT = 64; __VA_ARGS__; } \
        else if ((k) <= 100) { constexpr int KT = 100; __VA_ARGS__; } \
        else { constexpr int KT = 128; __VA_ARGS__; }            \
    } while (0)

// extra list slots for candidates that tie with the k-th neighbour in fp32
#define PCT_TIE_SLACK 16
