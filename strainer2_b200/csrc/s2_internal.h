// s2_internal.h - declarations shared between the translation units of libstrainer2_b200.so
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

void s2_set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

// environment knobs (argv of the drop-in executables stays identical to the reference's)
int      s2_env_int(const char *name, int dflt);
uint64_t s2_env_u64(const char *name, uint64_t dflt);
