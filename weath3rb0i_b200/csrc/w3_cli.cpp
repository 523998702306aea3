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
 * anything but book1 writes the block container, which carries the model),
 * `-w <bytes>` window: inputs (c) or outputs (d) larger than this are processed
 * a window of whole blocks at a time, so a file need not fit in host memory —
 * the container on disk is the same, byte for byte, as without it.
 */
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
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
          "Usage: w3 [-b block_bytes] [-g gpus] [-m book1|enwik7|train] [-w window_bytes] [-r 1] <action> <path>\n"
          "  action: c (compress) | d (decompress) | t (test)\n"
          "  -r 1: always write the reference's own file format (w30i: one stream = one block, at most 256 MiB);\n"
          "        without it files above 1 MiB go into the w3Bk block container, which the reference cannot read\n");
  exit(1);
}

static size_t g_block = 0;
static bool g_reference_format = false; /* -r 1 */
static size_t g_window = 0; /* 0: whole file in memory */
static std::string g_model = "book1";

static void be32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (24 - 8 * i)); }
static void be64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i)); }
static uint32_t rd32(const uint8_t *p) { uint32_t v = 0; for (int i = 0; i < 4; i++) v = (v << 8) | p[i]; return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | p[i]; return v; }
static const size_t HDR = 24, HDR_MODEL = 16; /* w3Bk container, include/w3_codec.h */

/* Windowed compress: the container's header and length index are laid out from
 * the file size alone, each window of whole blocks goes through w3_compress as a
 * container of its own, whose index entries and streams are copied into place. */
static int compress_windowed(w3_ctx *ctx, const std::string &in, const std::string &out, size_t n, size_t bs) {
  uint16_t table[8];
  int rc = w3_ctx_get_model(ctx, table);
  if (rc != W3_OK) return fail("model", rc);
  static const uint16_t book1[8] = {1, 50188, 62497, 15819, 22545, 31499, 22988, 29616}; /* stationary.rs:40-42 */
  const bool custom = memcmp(table, book1, sizeof book1) != 0;
  const size_t nb = (n + bs - 1) / bs, fixed = HDR + (custom ? HDR_MODEL : 0);
  const size_t wblocks = g_window / bs ? g_window / bs : 1;
  FILE *fi = fopen(in.c_str(), "rb"), *fo = fopen(out.c_str(), "wb");
  if (!fi || !fo) { perror(fi ? out.c_str() : in.c_str()); if (fi) fclose(fi); if (fo) fclose(fo); return 1; }
  std::vector<uint8_t> head(fixed + 4 * nb, 0);
  memcpy(head.data(), "w3Bk", 4);
  head[4] = custom ? 2 : 1;
  be64(head.data() + 8, n);
  be32(head.data() + 16, (uint32_t)bs);
  be32(head.data() + 20, (uint32_t)nb);
  if (custom) for (int i = 0; i < 8; i++) { head[HDR + 2 * i] = (uint8_t)(table[i] >> 8); head[HDR + 2 * i + 1] = (uint8_t)table[i]; }
  bool ok = fwrite(head.data(), 1, head.size(), fo) == head.size(); /* index filled in at the end */
  std::vector<uint8_t> win(wblocks * bs), buf(win.size() + win.size() / 4 + 4096);
  for (size_t b0 = 0; ok && b0 < nb; b0 += wblocks) {
    const size_t nbw = std::min(wblocks, nb - b0), bytes = std::min(n - b0 * bs, nbw * bs);
    if (fread(win.data(), 1, bytes, fi) != bytes) { ok = false; break; }
    size_t len = 0;
    rc = w3_compress(ctx, win.data(), bytes, bs, 0, buf.data(), buf.size(), &len);
    if (rc == W3_E_CAP) { buf.resize(len); rc = w3_compress(ctx, win.data(), bytes, bs, 0, buf.data(), buf.size(), &len); }
    if (rc != W3_OK) { fclose(fi); fclose(fo); return fail("compress", rc); }
    const size_t wfixed = HDR + (buf[4] == 2 ? HDR_MODEL : 0);
    memcpy(head.data() + fixed + 4 * b0, buf.data() + wfixed, 4 * nbw);
    const size_t soff = wfixed + 4 * nbw;
    ok = fwrite(buf.data() + soff, 1, len - soff, fo) == len - soff;
  }
  ok = ok && !fseek(fo, 0, SEEK_SET) && fwrite(head.data(), 1, head.size(), fo) == head.size();
  fclose(fi);
  ok = !fclose(fo) && ok;
  if (!ok) { perror(out.c_str()); return 1; }
  return 0;
}

/* Windowed decompress of a w3Bk container: header and index are read first,
 * then each window of blocks is handed to w3_decompress as a container of its own. */
static int decompress_windowed(w3_ctx *ctx, FILE *fi, const std::string &out, const uint8_t *h24) {
  const size_t n = (size_t)rd64(h24 + 8), bs = rd32(h24 + 16), nb = rd32(h24 + 20);
  const bool custom = h24[4] == 2;
  if ((h24[4] != 1 && h24[4] != 2) || !bs || nb != (n + bs - 1) / bs) return fail("decompress", W3_E_FORMAT);
  std::vector<uint8_t> model(custom ? HDR_MODEL : 0), index(4 * nb);
  if ((custom && fread(model.data(), 1, model.size(), fi) != model.size()) || fread(index.data(), 1, index.size(), fi) != index.size())
    return fail("decompress", W3_E_FORMAT);
  FILE *fo = fopen(out.c_str(), "wb");
  if (!fo) { perror(out.c_str()); return 1; }
  const size_t wblocks = g_window / bs ? g_window / bs : 1;
  std::vector<uint8_t> blob, data(wblocks * bs);
  bool ok = true;
  for (size_t b0 = 0; ok && b0 < nb; b0 += wblocks) {
    const size_t nbw = std::min(wblocks, nb - b0), bytes = std::min(n - b0 * bs, nbw * bs);
    size_t slen = 0;
    for (size_t b = b0; b < b0 + nbw; b++) slen += rd32(index.data() + 4 * b);
    const size_t wfixed = HDR + model.size();
    blob.assign(wfixed + 4 * nbw + slen, 0);
    memcpy(blob.data(), h24, HDR);
    be64(blob.data() + 8, bytes);
    be32(blob.data() + 20, (uint32_t)nbw);
    if (custom) memcpy(blob.data() + HDR, model.data(), HDR_MODEL);
    memcpy(blob.data() + wfixed, index.data() + 4 * b0, 4 * nbw);
    if (fread(blob.data() + wfixed + 4 * nbw, 1, slen, fi) != slen) { fclose(fo); return fail("decompress", W3_E_FORMAT); }
    size_t got = 0;
    const int rc = w3_decompress(ctx, blob.data(), blob.size(), data.data(), data.size(), &got);
    if (rc != W3_OK || got != bytes) { fclose(fo); return fail("decompress", rc != W3_OK ? rc : W3_E_FORMAT); }
    ok = fwrite(data.data(), 1, bytes, fo) == bytes;
  }
  ok = !fclose(fo) && ok;
  if (!ok) { perror(out.c_str()); return 1; }
  return 0;
}

static size_t file_size(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) ? 0 : (size_t)st.st_size;
}

static int do_compress(w3_ctx *ctx, const std::string &in, const std::string &out) {
  if (g_window && file_size(in) > g_window) {
    if (g_model == "train") { fprintf(stderr, "w3: -m train needs the whole input in memory (no -w)\n"); return 1; }
    auto t0 = std::chrono::steady_clock::now();
    if (g_model == "enwik7") {
      const uint16_t table[8] = {752, 50314, 58928, 21421, 24680, 30788, 24297, 32530}; /* for_enwik7 — stationary.rs:44-46 */
      const int mrc = w3_ctx_set_model(ctx, table);
      if (mrc != W3_OK) return fail("model", mrc);
    }
    const int r = compress_windowed(ctx, in, out, file_size(in), g_block ? g_block : W3_DEFAULT_BLOCK);
    if (!r) printf("Compression took: %.6fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); /* main.rs:74 */
    return r;
  }
  std::vector<uint8_t> data;
  if (!read_file(in, data)) { perror(in.c_str()); return 1; }
  auto t0 = std::chrono::steady_clock::now();
  if (g_model != "book1") {
    uint16_t table[8] = {752, 50314, 58928, 21421, 24680, 30788, 24297, 32530}; /* for_enwik7 — stationary.rs:44-46 */
    int mrc = g_model == "train" ? w3_train_model(ctx, data.data(), data.size(), table) : W3_OK;
    if (mrc == W3_OK) mrc = w3_ctx_set_model(ctx, table);
    if (mrc != W3_OK) return fail("model", mrc);
  }
  bool legacy = g_model == "book1" && g_block == 0 && (g_reference_format || data.size() <= W3_DEFAULT_BLOCK);
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
  if (g_window) { /* a block container whose output exceeds the window: block windows */
    FILE *fi = fopen(in.c_str(), "rb");
    uint8_t h[HDR];
    if (fi && fread(h, 1, HDR, fi) == HDR && !memcmp(h, "w3Bk", 4) && rd64(h + 8) > g_window) {
      auto t0 = std::chrono::steady_clock::now();
      const int r = decompress_windowed(ctx, fi, out, h);
      fclose(fi);
      if (!r) printf("Decompression took: %.6fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); /* main.rs:78 */
      return r;
    }
    if (fi) fclose(fi);
  }
  std::vector<uint8_t> blob;
  if (!read_file(in, blob)) { perror(in.c_str()); return 1; }
  auto t0 = std::chrono::steady_clock::now();
  size_t n = 0;
  int rc = w3_decompressed_size(blob.data(), blob.size(), &n);
  if (rc != W3_OK) return fail("decompress", rc);
  /* the length comes from an untrusted header: a coded bit costs at least log2(65536/65535) bits
   * (p <= 65535/65536), so a stream of s bytes cannot hold more than ~45427 * 8 * s bits */
  if (n / 45427 > blob.size() + 64) { fprintf(stderr, "w3: %s: the header claims %zu bytes for a %zu-byte file\n", in.c_str(), n, blob.size()); return 1; }
  std::vector<uint8_t> data;
  try { data.resize(n ? n : 1); } catch (const std::exception &) { fprintf(stderr, "w3: %s: cannot allocate %zu bytes\n", in.c_str(), n); return 1; }
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
    else if (!strcmp(argv[i], "-w")) g_window = strtoull(argv[i + 1], nullptr, 0);
    else if (!strcmp(argv[i], "-r")) g_reference_format = atoi(argv[i + 1]) != 0;
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
