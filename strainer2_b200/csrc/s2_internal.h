// s2_internal.h - declarations shared between the translation units of libstrainer2_b200.so
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

bool s2_list_starts_with_plain_gz(const char *list_file);      // s2_cli_count.cpp

void s2_set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

// environment knobs (argv of the drop-in executables stays identical to the reference's)
int      s2_env_int(const char *name, int dflt);
uint64_t s2_env_u64(const char *name, uint64_t dflt);

// ---- host string path for windows containing bytes outside ACGTN (s2_exotic.cpp) ----------
struct s2_exotic;
struct S2ExoRow { std::string key; uint64_t first_pos; uint32_t djb2; uint32_t counts[8]; };
s2_exotic *s2_exotic_build(const uint8_t *flat, uint64_t n, int n_cols);     // nullptr when there is nothing to do
void       s2_exotic_free(s2_exotic *ex);
uint64_t   s2_exotic_n_keys(const s2_exotic *ex);
uint64_t   s2_exotic_n_informative(const s2_exotic *ex);
void       s2_exotic_count_record(s2_exotic *ex, const char *seq, uint64_t len, int col);
void       s2_exotic_rows(const s2_exotic *ex, std::vector<S2ExoRow> &rows);
bool       s2_exotic_flag(s2_exotic *ex, const char *line31);
bool       s2_exotic_is_informative(const s2_exotic *ex, const char *key);
void       s2_exotic_set_informative(s2_exotic *ex, const char *key, bool v);
void       s2_exotic_pass1(s2_exotic *ex, const char *seq, uint64_t len, int *hits, int *inf);
void       s2_exotic_pass2(s2_exotic *ex, const char *seq, uint64_t len, std::vector<std::pair<uint64_t, std::string>> &out);

// ---- shared by the two executables (s2_cli_count.cpp) --------------------------------------------------
struct s2_ctx; struct s2_table;
struct S2WorkItem { std::string path; int col; bool skip; };
int  s2_read_list(const char *list_file, int col, const char *skip_file, std::vector<S2WorkItem> &out);
bool s2_scan_work_items(s2_ctx *ctx, s2_table *table, s2_exotic *exotic, std::vector<S2WorkItem> &work, int n_threads,
                        FILE *progress, std::string &open_error, uint64_t *bases_out, uint64_t *lookups_out);
bool s2_scan_work_items_multi(std::vector<s2_ctx *> &ctxs, std::vector<s2_table *> &tables, s2_exotic *exotic,
                              std::vector<S2WorkItem> &work, int n_threads, FILE *progress, std::string &open_error,
                              uint64_t *bases_out, uint64_t *lookups_out);
int  s2_default_reader_threads();
