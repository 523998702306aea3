/*
 * w3_corpus.c — fixed-seed synthetic English-like text.
 *
 * Bench/test input generator, not part of the codec.  Integer-only so that the
 * Python twin (corpus/w3_corpus.py) produces identical bytes.  The stream is a
 * concatenation of independent 64 KiB chunks, chunk k seeded from
 * SEED ^ (k * GOLDEN), so any prefix of any length is stable and chunks can be
 * generated in parallel.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define W3C_SEED 0x77333069ull
#define GOLDEN 0x9E3779B97F4A7C15ull
#define CHUNK 65536u
#define VOCAB 8192
#define MAXW 12

static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += GOLDEN);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static const char *const FUNCTION_WORDS[64] = {
    "the",  "of",   "and",   "to",    "a",     "in",    "that", "is",
    "was",  "he",   "for",   "it",    "with",  "as",    "his",  "on",
    "be",   "at",   "by",    "i",     "this",  "had",   "not",  "are",
    "but",  "from", "or",    "have",  "an",    "they",  "which", "one",
    "you",  "were", "her",   "all",   "she",   "there", "would", "their",
    "we",   "him",  "been",  "has",   "when",  "who",   "will", "more",
    "no",   "if",   "out",   "so",    "said",  "what",  "up",   "its",
    "about", "into", "than", "them",  "can",   "only",  "other", "new"};

/* word-length weights for lengths 1..12 (sum 256) */
static const uint16_t LEN_W[MAXW] = {2, 14, 34, 46, 44, 38, 30, 20, 13, 8, 4, 3};
/* letter weights (sum 1024 each) */
static const char VOWELS[] = "eaoiuy";
static const uint16_t VOWEL_W[6] = {330, 220, 200, 190, 70, 14};
static const char CONS[] = "tnshrdlcmwfgpbvkjxqz";
static const uint16_t CONS_W[20] = {150, 112, 105, 101, 99, 71, 67, 46, 40, 39,
                                    37,  33,  32,  25,  16, 13, 3,  3,  2,  30};

typedef struct {
  char words[VOCAB][MAXW + 1];
  uint8_t wlen[VOCAB];
  uint64_t cum[VOCAB]; /* cumulative Zipf weights, w_i = 2^32 / (i+1) */
  uint64_t total;
} vocab_t;

static vocab_t *g_vocab;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static unsigned pick(uint64_t r, const uint16_t *w, int n, unsigned sum) {
  unsigned x = (unsigned)(r % sum);
  for (int i = 0; i < n; i++) {
    if (x < w[i]) return (unsigned)i;
    x -= w[i];
  }
  return (unsigned)(n - 1);
}

static void build_vocab(void) {
  vocab_t *v = (vocab_t *)calloc(1, sizeof *v);
  uint64_t s = W3C_SEED;
  for (int i = 0; i < VOCAB; i++) {
    if (i < 64) {
      size_t L = strlen(FUNCTION_WORDS[i]);
      memcpy(v->words[i], FUNCTION_WORDS[i], L);
      v->wlen[i] = (uint8_t)L;
    } else {
      unsigned L = 1 + pick(splitmix64(&s) >> 16, LEN_W, MAXW, 256);
      if (L < 2) L = 2;
      int vowel = (int)((splitmix64(&s) >> 20) & 3) == 0; /* start: 1/4 vowel */
      for (unsigned k = 0; k < L; k++) {
        uint64_t r = splitmix64(&s) >> 16;
        if (vowel) v->words[i][k] = VOWELS[pick(r, VOWEL_W, 6, 1024)];
        else v->words[i][k] = CONS[pick(r, CONS_W, 20, 1024)];
        /* alternation bias: after a vowel 7/8 consonant; after a consonant 3/4 vowel */
        unsigned t = (unsigned)((r >> 24) & 7);
        vowel = vowel ? (t == 0) : (t < 6);
      }
      v->wlen[i] = (uint8_t)L;
    }
    v->total += ((uint64_t)1 << 32) / (uint64_t)(i + 1);
    v->cum[i] = v->total;
  }
  g_vocab = v;
}

static inline unsigned zipf_rank(const vocab_t *v, uint64_t r) {
  uint64_t x = (r >> 11) % v->total;
  unsigned lo = 0, hi = VOCAB - 1;
  while (lo < hi) {
    unsigned mid = (lo + hi) >> 1;
    if (x < v->cum[mid]) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

/* Generates chunk k (exactly CHUNK bytes) into out. */
static void gen_chunk(uint64_t k, uint8_t *out) {
  const vocab_t *v = g_vocab;
  uint64_t s = W3C_SEED ^ (k * GOLDEN);
  uint8_t buf[CHUNK + 64];
  size_t n = 0;
  unsigned recent[8];
  unsigned n_recent = 0;
  int sentence_start = 1;
  while (n < CHUNK) {
    uint64_t r = splitmix64(&s);
    if (((r >> 8) & 63) == 0) { /* rare digit group */
      unsigned nd = 1 + (unsigned)((r >> 16) & 3);
      uint64_t d = r >> 20;
      for (unsigned i = 0; i < nd; i++) {
        buf[n++] = (uint8_t)('0' + d % 10);
        d /= 10;
      }
    } else {
      unsigned w;
      if (n_recent == 8 && ((r >> 16) & 15) == 0) w = recent[(r >> 20) & 7];
      else w = zipf_rank(v, splitmix64(&s));
      recent[n_recent < 8 ? n_recent++ : (unsigned)((r >> 24) & 7)] = w;
      memcpy(buf + n, v->words[w], v->wlen[w]);
      if (sentence_start) buf[n] = (uint8_t)(buf[n] - 32);
      n += v->wlen[w];
    }
    sentence_start = 0;
    unsigned e = (unsigned)((r >> 32) % 168); /* lcm-ish of 14 and 12 */
    if (e < 12) { /* sentence end, 1/14 */
      buf[n++] = '.';
      buf[n++] = (((r >> 44) % 5) == 0) ? '\n' : ' ';
      sentence_start = 1;
    } else if (e < 26) { /* comma, 1/12 */
      buf[n++] = ',';
      buf[n++] = ' ';
    } else {
      buf[n++] = ' ';
    }
  }
  memcpy(out, buf, CHUNK);
}

typedef struct {
  uint8_t *out;
  uint64_t first_chunk, n_chunks, offset, n;
  uint64_t *next;
  pthread_mutex_t *mu;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  uint8_t tmp[CHUNK];
  for (;;) {
    pthread_mutex_lock(j->mu);
    uint64_t c = (*j->next)++;
    pthread_mutex_unlock(j->mu);
    if (c >= j->n_chunks) break;
    uint64_t k = j->first_chunk + c;
    uint64_t lo = k * CHUNK, hi = lo + CHUNK; /* absolute byte range */
    uint64_t a = lo > j->offset ? lo : j->offset;
    uint64_t b = hi < j->offset + j->n ? hi : j->offset + j->n;
    if (a == lo && b == hi) {
      gen_chunk(k, j->out + (lo - j->offset));
    } else {
      gen_chunk(k, tmp);
      memcpy(j->out + (a - j->offset), tmp + (a - lo), (size_t)(b - a));
    }
  }
  return NULL;
}

/* Fills out[0..n) with bytes [offset, offset+n) of the infinite corpus stream. */
int w3c_generate(uint8_t *out, uint64_t offset, uint64_t n, int n_threads) {
  if (!out && n) return -1;
  if (n == 0) return 0;
  pthread_once(&g_once, build_vocab);
  job_t proto;
  proto.out = out;
  proto.offset = offset;
  proto.n = n;
  proto.first_chunk = offset / CHUNK;
  proto.n_chunks = (offset + n + CHUNK - 1) / CHUNK - proto.first_chunk;
  if (n_threads < 1) n_threads = 1;
  if ((uint64_t)n_threads > proto.n_chunks) n_threads = (int)proto.n_chunks;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  uint64_t next = 0;
  proto.next = &next;
  proto.mu = &mu;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)n_threads);
  for (int i = 0; i < n_threads; i++) {
    jobs[i] = proto;
    pthread_create(&th[i], NULL, worker, &jobs[i]);
  }
  for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
  free(th);
  free(jobs);
  return 0;
}
