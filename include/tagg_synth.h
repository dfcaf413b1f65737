/*
 * tagg_synth.h — on-device synthetic column generators (bench / scale tests only).
 *
 * Not part of the reference's surface: the reference's bench builds its corpus with a host RNG
 * (benches/lib.rs:70-90).  At 10^8..10^9 docs that is impractical, so columns are generated in
 * HBM with the counter-based recipe of SURVEY §8d; oracle/oracle.cpp restates the same recipe
 * on the CPU (orc_synth_codes / orc_synth_multi) so samples can be checked bit for bit.
 *
 *   x(doc)   = mix64(seed ^ tag ^ (doc_base + doc) * 0x9E3779B97F4A7C15)      (splitmix64 finaliser)
 *   recipe 0 : code(f64 1.0 + 100.0 * ((x >> 11) * 2^-53))   price in [1, 101)
 *   recipe 1 : a + x mod b
 *   recipe 2 : a + (x mod b) * c                              b distinct keys over a wide domain
 *   recipe 3 : a + floor(b * u^4), u = (x >> 32) * 2^-32      power-law (Zipf-like) keys in [a, a + b), b < 2^32
 * multi-valued: count(doc) = x(doc, tag ^ 0xC0FFEE1234567) mod count_mod,
 *               value j    = recipe(mix64(x(doc, tag) + (j+1) * 0xD6E8FEB86659FD93))
 * The generated codes go through the same device pack path as tagg_column_upload_codes.
 */
#ifndef TAGG_SYNTH_H
#define TAGG_SYNTH_H
#include "tagg.h"
#ifdef __cplusplus
extern "C" {
#endif
int tagg_synth_column(tagg_segment* seg, uint32_t field_id, int kind, int recipe, uint64_t seed,
                      uint64_t tag, uint64_t doc_base, uint64_t a, uint64_t b, uint64_t c);
int tagg_synth_multicolumn(tagg_segment* seg, uint32_t field_id, int kind, int recipe, uint64_t seed,
                           uint64_t tag, uint64_t doc_base, uint64_t count_mod, uint64_t a, uint64_t b,
                           uint64_t c);
#ifdef __cplusplus
}
#endif
#endif
