/*
 * w3 — command-line driver over libw3codec.so, mirroring the reference binary
 * `weath3rb0i <c|d|t> <path>` (src/main.rs:24-87, usage 154-161): outputs go to
 * the current directory as <stem>.bin (c, t) or <stem>.orig (d); a directory
 * argument is walked one level.  `t` also compares the round trip with the
 * input, which the reference leaves to helpers::cmp (src/helpers.rs:7-28).
 *
 * Extra, not in the reference: `-b <bytes>` block size (default: the
 * reference's single-stream file when the input fits 1 MiB, else 1 MiB blocks),
 * `-g <n>` number of GPUs, `-m book1|enwik7|train` the StationaryModel of the
 * entropy hash (src/models/ac_hash/stationary.rs: for_book1 is what main.rs
 * uses; `train` = StationaryModel::new on the input, as the sweep binaries do;
 * anything but book1 writes the block container, which carries the model).
 */
#include <dirent.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/w3_codec.h"

static bool read_file(const std::string &p, std::vector<uint8_t> &v) {
  FILE *f = fopen(p.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize((size_t)n);
  bool ok = n == 0 || fread(v.data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  return ok;
}
static bool write_file(const std::string &p, const uint8_t *d, size_t n) {
  FILE *f = fopen(p.c_str(), "wb");
  if (!f) return false;
  bool ok = n == 0 || fwrite(d, 1, n, f) == n;
  fclose(f);
  return ok;
}
static std::string stem_of(const std::string &path) {
  size_t s = path.find_last_of('/');
  std::string name = s == std::string::npos ? path : path.substr(s + 1);
  size_t d = name.find_last_of('.');
  return d == std::string::npos || d == 0 ? name : name.substr(0, d);
}
static int fail(const char *what, int rc) {
  fprintf(stderr, "%s: %s%s%s\n", what, w3_strerror(rc), rc == W3_E_CUDA ? " — " : "",
          rc == W3_E_CUDA ? w3_last_cuda_error() : "");
  return 1;
}
static void usage() { /* main.rs:154-161 */
  fprintf(stderr,
          "Usage: w3 [-b block_bytes] [-g gpus] [-m book1|enwik7|train] <action> <path>\n"
          "  action: c (compress) | d (decompress) | t (test)\n");
  exit(1);
}

static size_t g_block = 0;
static std::string g_model = "book1";

static int do_compress(w3_ctx *ctx, const std::string &in, const std::string &out) {
  std::vector<uint8_t> data;
  if (!read_file(in, data)) { perror(in.c_str()); return 1; }
  auto t0 = std::chrono::steady_clock::now();
  if (g_model != "book1") {
    uint16_t table[8] = {752, 50314, 58928, 21421, 24680, 30788, 24297, 32530}; /* for_enwik7 — stationary.rs:44-46 */
    int mrc = g_model == "train" ? w3_train_model(ctx, data.data(), data.size(), table) : W3_OK;
    if (mrc == W3_OK) mrc = w3_ctx_set_model(ctx, table);
    if (mrc != W3_OK) return fail("model", mrc);
  }
  bool legacy = g_model == "book1" && g_block == 0 && data.size() <= W3_DEFAULT_BLOCK;
  size_t bs = g_block ? g_block : W3_DEFAULT_BLOCK;
  std::vector<uint8_t> buf(data.size() + data.size() / 4 + 4096);
  size_t len = 0;
  int rc = w3_compress(ctx, data.data(), data.size(), bs, legacy ? W3_F_LEGACY : 0, buf.data(), buf.size(), &len);
  if (rc == W3_E_CAP) {
    buf.resize(len);
    rc = w3_compress(ctx, data.data(), data.size(), bs, legacy ? W3_F_LEGACY : 0, buf.data(), buf.size(), &len);
  }
  if (rc != W3_OK) return fail("compress", rc);
  if (!write_file(out, buf.data(), len)) { perror(out.c_str()); return 1; }
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("Compression took: %.6fs\n", s); /* main.rs:74 */
  return 0;
}
static int do_decompress(w3_ctx *ctx, const std::string &in, const std::string &out) {
  std::vector<uint8_t> blob;
  if (!read_file(in, blob)) { perror(in.c_str()); return 1; }
  auto t0 = std::chrono::steady_clock::now();
  size_t n = 0;
  int rc = w3_decompressed_size(blob.data(), blob.size(), &n);
  if (rc != W3_OK) return fail("decompress", rc);
  std::vector<uint8_t> data(n ? n : 1);
  rc = w3_decompress(ctx, blob.data(), blob.size(), data.data(), n, &n);
  if (rc != W3_OK) return fail("decompress", rc);
  if (!write_file(out, data.data(), n)) { perror(out.c_str()); return 1; }
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("Decompression took: %.6fs\n", s); /* main.rs:78 */
  return 0;
}
static int run(w3_ctx *ctx, const std::string &path, char action) { /* main.rs:55-87 */
  std::string bin = stem_of(path) + ".bin", orig = stem_of(path) + ".orig";
  if (action == 'c') return do_compress(ctx, path, bin);
  if (action == 'd') return do_decompress(ctx, path, orig);
  if (do_compress(ctx, path, bin) || do_decompress(ctx, bin, orig)) return 1;
  std::vector<uint8_t> a, b;
  bool same = read_file(path, a) && read_file(orig, b) && a == b;
  printf("round trip: %s\n", same ? "identical" : "MISMATCH");
  return same ? 0 : 2;
}

int main(int argc, char **argv) {
  int gpus = 1, i = 1;
  for (; i + 1 < argc && argv[i][0] == '-'; i += 2) {
    if (!strcmp(argv[i], "-b")) g_block = strtoull(argv[i + 1], nullptr, 0);
    else if (!strcmp(argv[i], "-g")) gpus = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "-m") && (!strcmp(argv[i + 1], "book1") || !strcmp(argv[i + 1], "enwik7") || !strcmp(argv[i + 1], "train"))) g_model = argv[i + 1];
    else usage();
  }
  if (argc - i != 2 || strlen(argv[i]) != 1 || !strchr("cdt", argv[i][0])) usage();
  char action = argv[i][0];
  std::string path = argv[i + 1];
  w3_ctx *ctx = nullptr;
  int rc = w3_ctx_create(nullptr, gpus, &ctx);
  if (rc != W3_OK) return fail("w3_ctx_create", rc);
  struct stat st;
  if (stat(path.c_str(), &st)) { perror(path.c_str()); return 1; }
  int ret = 0;
  if (S_ISDIR(st.st_mode)) { /* main.rs:41-47 */
    DIR *d = opendir(path.c_str());
    std::vector<std::string> names;
    for (dirent *e; d && (e = readdir(d));) names.push_back(e->d_name);
    if (d) closedir(d);
    for (auto &nm : names) {
      std::string p = path + "/" + nm;
      if (!stat(p.c_str(), &st) && S_ISREG(st.st_mode)) ret |= run(ctx, p, action);
    }
  } else {
    ret = run(ctx, path, action);
  }
  w3_ctx_destroy(ctx);
  return ret;
}
