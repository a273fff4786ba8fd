/* TEST INFRASTRUCTURE ONLY -- minimal FITS BINTABLE writer behind the cfitsio names the reference's
 * he_write_healpix_map calls (healpix_extra.c:132-164), and a small reader behind the ones he_read_healpix_map calls
 * (healpix_extra.c:166-224).  Like cfitsio, fits_create_file fails when the file already exists (status 105) and later
 * calls are no-ops once status is non-zero. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fitsio.h>

#define MAXCARDS 128
#define MAXCOLS 3
struct shim_fitsfile {
  FILE *fp;
  int ncols;
  char ttype[MAXCOLS][16], tform[MAXCOLS][16], tunit[MAXCOLS][16], extname[32];
  char cards[MAXCARDS][81];
  int ncards;
  float *col[MAXCOLS];
  long nrows;
};

static void card(char *dst, const char *key, const char *val, const char *comm)
{
  char buf[256];
  if (comm && *comm) snprintf(buf, sizeof(buf), "%-8.8s= %20s / %s", key, val, comm);
  else snprintf(buf, sizeof(buf), "%-8.8s= %20s", key, val);
  snprintf(dst, 81, "%-80.80s", buf);
}
static void card_str(char *dst, const char *key, const char *val, const char *comm)
{
  char q[96], buf[256];
  snprintf(q, sizeof(q), "'%-8s'", val);
  if (comm && *comm) snprintf(buf, sizeof(buf), "%-8.8s= %-20s / %s", key, q, comm);
  else snprintf(buf, sizeof(buf), "%-8.8s= %-20s", key, q);
  snprintf(dst, 81, "%-80.80s", buf);
}
static void put_block(FILE *fp, char (*cards)[81], int n)
{
  long bytes = 0;
  for (int i = 0; i < n; i++) { fwrite(cards[i], 1, 80, fp); bytes += 80; }
  char end[81];
  snprintf(end, 81, "%-80s", "END");
  fwrite(end, 1, 80, fp); bytes += 80;
  while (bytes % 2880) { fputc(' ', fp); bytes++; }
}

int fits_create_file(fitsfile **fptr, const char *filename, int *status)
{
  *fptr = NULL;
  if (*status) return *status;
  FILE *t = fopen(filename, "rb");
  if (t) { fclose(t); return (*status = 105); }
  FILE *fp = fopen(filename, "wb");
  if (!fp) return (*status = 105);
  fitsfile *f = calloc(1, sizeof(*f));
  f->fp = fp;
  *fptr = f;
  return 0;
}
int fits_create_tbl(fitsfile *f, int tbltype, long naxis2, int tfields, char **ttype, char **tform, char **tunit,
                    const char *extname, int *status)
{
  (void)tbltype; (void)naxis2;
  if (*status || !f) return *status;
  f->ncols = tfields;
  for (int i = 0; i < tfields && i < MAXCOLS; i++) {
    snprintf(f->ttype[i], 16, "%s", ttype[i]);
    snprintf(f->tform[i], 16, "%s", tform[i]);
    snprintf(f->tunit[i], 16, "%s", tunit[i]);
  }
  snprintf(f->extname, 32, "%s", extname);
  return 0;
}
int fits_write_key(fitsfile *f, int datatype, const char *keyname, void *value, const char *comm, int *status)
{
  if (*status || !f) return *status;
  if (f->ncards >= MAXCARDS) return (*status = 1);
  if (datatype == TSTRING) card_str(f->cards[f->ncards++], keyname, (const char *)value, comm);
  else if (datatype == TLONG) {
    char v[32];
    snprintf(v, 32, "%ld", *(long *)value);
    card(f->cards[f->ncards++], keyname, v, comm);
  } else return (*status = 1);
  return 0;
}
int fits_write_comment(fitsfile *f, const char *comm, int *status)
{
  if (*status || !f) return *status;
  char buf[128];
  snprintf(buf, sizeof(buf), "COMMENT %s", comm);
  snprintf(f->cards[f->ncards++], 81, "%-80.80s", buf);
  return 0;
}
int fits_write_col(fitsfile *f, int datatype, int colnum, long firstrow, long firstelem, long nelem, void *array,
                   int *status)
{
  if (*status || !f) return *status;
  if (datatype != TFLOAT || firstrow != 1 || firstelem != 1 || colnum < 1 || colnum > f->ncols) return (*status = 1);
  f->col[colnum - 1] = malloc(sizeof(float) * nelem);
  memcpy(f->col[colnum - 1], array, sizeof(float) * nelem);
  f->nrows = nelem;
  return 0;
}
static int close_reader(fitsfile *f, int *status);
int fits_close_file(fitsfile *f, int *status)
{
  if (!f) return *status;
  if (!f->fp) return close_reader(f, status); /* a handle from fits_open_file */
  char c[MAXCARDS][81];
  int n = 0;
  char v[32];
  card(c[n++], "SIMPLE", "T", "file does conform to FITS standard");
  card(c[n++], "BITPIX", "8", "number of bits per data pixel");
  card(c[n++], "NAXIS", "0", "number of data axes");
  card(c[n++], "EXTEND", "T", "FITS dataset may contain extensions");
  put_block(f->fp, c, n);
  n = 0;
  card_str(c[n++], "XTENSION", "BINTABLE", "binary table extension");
  card(c[n++], "BITPIX", "8", "8-bit bytes");
  card(c[n++], "NAXIS", "2", "2-dimensional binary table");
  snprintf(v, 32, "%d", 4 * f->ncols); card(c[n++], "NAXIS1", v, "width of table in bytes");
  snprintf(v, 32, "%ld", f->nrows); card(c[n++], "NAXIS2", v, "number of rows in table");
  card(c[n++], "PCOUNT", "0", "size of special data area");
  card(c[n++], "GCOUNT", "1", "one data group (required keyword)");
  snprintf(v, 32, "%d", f->ncols); card(c[n++], "TFIELDS", v, "number of fields in each row");
  for (int i = 0; i < f->ncols; i++) {
    char k[16];
    snprintf(k, 16, "TTYPE%d", i + 1); card_str(c[n++], k, f->ttype[i], "label for field");
    snprintf(k, 16, "TFORM%d", i + 1); card_str(c[n++], k, f->tform[i], "data format of field: 4-byte REAL");
    snprintf(k, 16, "TUNIT%d", i + 1); card_str(c[n++], k, f->tunit[i], "physical unit of field");
  }
  card_str(c[n++], "EXTNAME", f->extname, "name of this binary table extension");
  for (int i = 0; i < f->ncards; i++) memcpy(c[n++], f->cards[i], 81);
  put_block(f->fp, c, n);
  long bytes = 0;
  for (long r = 0; r < f->nrows; r++)
    for (int i = 0; i < f->ncols; i++) {
      uint32_t u;
      memcpy(&u, &f->col[i][r], 4);
      unsigned char be[4] = {(unsigned char)(u >> 24), (unsigned char)(u >> 16), (unsigned char)(u >> 8), (unsigned char)u};
      fwrite(be, 1, 4, f->fp);
      bytes += 4;
    }
  while (bytes % 2880) { fputc(0, f->fp); bytes++; }
  fclose(f->fp);
  for (int i = 0; i < MAXCOLS; i++) free(f->col[i]);
  free(f);
  return *status;
}

/* ---- read path (he_read_healpix_map, healpix_extra.c:166-224; JoinT's inputs) --------------------------------
 * A small FITS reader written against the FITS standard, not against the writer above: any sequence of HDUs in
 * 2880-byte blocks, 80-character cards, BINTABLE columns of rE / rD (big-endian IEEE), row-major.  Like cfitsio,
 * every call is a no-op once *status is non-zero; a missing key sets status 202 (KEY_NO_EXIST), a missing file 104. */
struct hdu_info {
  long hdr_off, data_off, data_bytes;
  int ncards;
};
#define MAXHDU 8
struct shim_reader {
  FILE *fp;
  int nhdu, cur;
  struct hdu_info hdu[MAXHDU];
  char (*cards)[81]; /* cards of the current HDU */
  int ncards;
};
/* one handle type for both directions: a reader is recognised by fp == NULL in the writer part */
struct shim_rw {
  struct shim_fitsfile w; /* must stay first: the write path casts fitsfile* to this */
  struct shim_reader r;
};

static int card_key_is(const char *card, const char *key)
{
  char k[9];
  memcpy(k, card, 8);
  k[8] = 0;
  for (int i = 7; i >= 0 && k[i] == ' '; i--) k[i] = 0;
  return strcmp(k, key) == 0;
}
static const char *find_card(const struct shim_reader *r, const char *key)
{
  for (int i = 0; i < r->ncards; i++)
    if (card_key_is(r->cards[i], key) && r->cards[i][8] == '=') return r->cards[i];
  return NULL;
}
static int card_long(const struct shim_reader *r, const char *key, long *v)
{
  const char *c = find_card(r, key);
  if (!c) return 0;
  *v = strtol(c + 10, NULL, 10);
  return 1;
}
/* the quoted string value of a card, trailing blanks removed */
static int card_string(const struct shim_reader *r, const char *key, char *out, size_t cap)
{
  const char *c = find_card(r, key);
  if (!c) return 0;
  const char *q = strchr(c + 10, '\'');
  if (!q) return 0;
  q++;
  size_t n = 0;
  while (*q && n + 1 < cap) {
    if (*q == '\'') {
      if (q[1] == '\'') q++; /* doubled quote */
      else break;
    }
    out[n++] = *q++;
  }
  while (n > 0 && out[n - 1] == ' ') n--;
  out[n] = 0;
  return 1;
}
static int load_cards(struct shim_reader *r, int ihdu)
{
  const struct hdu_info *h = &r->hdu[ihdu];
  free(r->cards);
  r->cards = malloc((size_t)h->ncards * 81);
  if (!r->cards) return 0;
  fseek(r->fp, h->hdr_off, SEEK_SET);
  for (int i = 0; i < h->ncards; i++) {
    if (fread(r->cards[i], 1, 80, r->fp) != 80) return 0;
    r->cards[i][80] = 0;
  }
  r->ncards = h->ncards;
  r->cur = ihdu;
  return 1;
}
/* walk the file: header blocks up to END, then the data area rounded up to whole blocks */
static int scan_hdus(struct shim_reader *r)
{
  long off = 0;
  r->nhdu = 0;
  for (;;) {
    char card[81];
    struct hdu_info h = {off, 0, 0, 0};
    long bitpix = 8, naxis = 0, pcount = 0, gcount = 1, prod = 1;
    int end = 0, any = 0;
    fseek(r->fp, off, SEEK_SET);
    while (!end) {
      if (fread(card, 1, 80, r->fp) != 80) { if (any) return 0; return r->nhdu > 0; }
      any = 1;
      card[80] = 0;
      if (card_key_is(card, "END")) { end = 1; break; }
      h.ncards++;
      if (card[8] != '=') continue;
      if (card_key_is(card, "BITPIX")) bitpix = strtol(card + 10, NULL, 10);
      else if (card_key_is(card, "NAXIS")) naxis = strtol(card + 10, NULL, 10);
      else if (card_key_is(card, "PCOUNT")) pcount = strtol(card + 10, NULL, 10);
      else if (card_key_is(card, "GCOUNT")) gcount = strtol(card + 10, NULL, 10);
      else if (!strncmp(card, "NAXIS", 5) && card[5] >= '1' && card[5] <= '9') prod *= strtol(card + 10, NULL, 10);
    }
    long hdr_bytes = ((long)(h.ncards + 1) * 80 + 2879) / 2880 * 2880;
    h.data_off = off + hdr_bytes;
    long bytes_per = bitpix < 0 ? -bitpix / 8 : bitpix / 8;
    h.data_bytes = naxis ? bytes_per * gcount * (pcount + prod) : 0;
    if (r->nhdu == MAXHDU) return 0;
    r->hdu[r->nhdu++] = h;
    off = h.data_off + (h.data_bytes + 2879) / 2880 * 2880;
  }
}

int fits_open_file(fitsfile **fptr, const char *filename, int mode, int *status)
{
  (void)mode;
  *fptr = NULL;
  if (*status) return *status;
  FILE *fp = fopen(filename, "rb");
  if (!fp) return (*status = 104);
  struct shim_rw *f = calloc(1, sizeof(*f));
  f->r.fp = fp;
  if (!scan_hdus(&f->r) || !load_cards(&f->r, 0)) { fclose(fp); free(f); return (*status = 252); }
  *fptr = (fitsfile *)f;
  return 0;
}
#define READER(f) (&((struct shim_rw *)(f))->r)
static int close_reader(fitsfile *f, int *status)
{
  struct shim_reader *r = READER(f);
  if (r->fp) fclose(r->fp);
  free(r->cards);
  free(f);
  return *status;
}
int fits_movabs_hdu(fitsfile *fptr, int hdunum, int *exttype, int *status)
{
  if (*status) return *status;
  struct shim_reader *r = READER(fptr);
  if (hdunum < 1 || hdunum > r->nhdu || !load_cards(r, hdunum - 1)) return (*status = 107); /* END_OF_FILE */
  if (exttype) {
    char x[32] = "";
    *exttype = 0; /* IMAGE_HDU */
    if (card_string(r, "XTENSION", x, sizeof(x)) && !strcmp(x, "BINTABLE")) *exttype = BINARY_TBL;
  }
  return 0;
}
int fits_read_key_lng(fitsfile *fptr, const char *keyname, long *value, char *comm, int *status)
{
  (void)comm;
  if (*status) return *status;
  if (!card_long(READER(fptr), keyname, value)) return (*status = 202);
  return 0;
}
int fits_read_keys_lng(fitsfile *fptr, const char *keyname, int nstart, int nmax, long *value, int *nfound, int *status)
{
  if (*status) return *status;
  *nfound = 0;
  for (int i = 0; i < nmax; i++) {
    char k[16];
    snprintf(k, sizeof(k), "%s%d", keyname, nstart + i);
    if (card_long(READER(fptr), k, &value[i])) (*nfound)++;
  }
  return 0;
}
int fits_read_key(fitsfile *fptr, int datatype, const char *keyname, void *value, char *comm, int *status)
{
  (void)comm;
  if (*status) return *status;
  if (datatype == TSTRING) {
    if (!card_string(READER(fptr), keyname, (char *)value, 32)) return (*status = 202);
    return 0;
  }
  if (datatype == TLONG) {
    if (!card_long(READER(fptr), keyname, (long *)value)) return (*status = 202);
    return 0;
  }
  return (*status = 410); /* BAD_DATATYPE */
}
/* TFORMn = rE or rD (r optional): repeat count and element size */
static int parse_tform(const char *t, long *rep, int *size, char *code)
{
  char *e;
  long r = strtol(t, &e, 10);
  if (e == t) r = 1;
  if (*e != 'E' && *e != 'D') return 0;
  *rep = r;
  *size = (*e == 'E') ? 4 : 8;
  *code = *e;
  return 1;
}
int fits_read_col(fitsfile *fptr, int datatype, int colnum, long firstrow, long firstelem, long nelem, void *nulval,
                  void *array, int *anynul, int *status)
{
  (void)nulval;
  if (*status) return *status;
  struct shim_reader *r = READER(fptr);
  long tfields = 0, naxis1 = 0, naxis2 = 0;
  if (!card_long(r, "TFIELDS", &tfields) || !card_long(r, "NAXIS1", &naxis1) || !card_long(r, "NAXIS2", &naxis2) || colnum < 1 ||
      colnum > tfields)
    return (*status = 302); /* BAD_COL_NUM */
  long col_off = 0, rep = 0;
  int size = 0;
  char code = 0;
  for (int c = 1; c <= colnum; c++) {
    char k[16], t[32];
    long rc;
    int sc;
    char cc;
    snprintf(k, sizeof(k), "TFORM%d", c);
    if (!card_string(r, k, t, sizeof(t)) || !parse_tform(t, &rc, &sc, &cc)) return (*status = 261); /* BAD_TFORM */
    if (c == colnum) { rep = rc; size = sc; code = cc; }
    else col_off += rc * sc;
  }
  if (anynul) *anynul = 0;
  long elem = (firstrow - 1) * rep + (firstelem - 1);
  if (elem < 0 || elem + nelem > naxis2 * rep) return (*status = 307); /* BAD_ROW_NUM */
  unsigned char *rowbuf = malloc((size_t)rep * size);
  long done = 0;
  while (done < nelem) {
    const long row = (elem + done) / rep, e0 = (elem + done) % rep;
    long n = rep - e0;
    if (n > nelem - done) n = nelem - done;
    fseek(r->fp, r->hdu[r->cur].data_off + row * naxis1 + col_off + e0 * size, SEEK_SET);
    if (fread(rowbuf, (size_t)size, (size_t)n, r->fp) != (size_t)n) { free(rowbuf); return (*status = 108); /* READ_ERROR */ }
    for (long i = 0; i < n; i++) {
      const unsigned char *b = rowbuf + i * size;
      double v;
      if (code == 'E') {
        uint32_t u = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
        float f;
        memcpy(&f, &u, 4);
        v = f;
        if (datatype == TFLOAT) { ((float *)array)[done + i] = f; continue; }
      } else {
        uint64_t u = 0;
        for (int j = 0; j < 8; j++) u = (u << 8) | b[j];
        memcpy(&v, &u, 8);
      }
      if (datatype == TFLOAT) ((float *)array)[done + i] = (float)v;
      else if (datatype == TDOUBLE) ((double *)array)[done + i] = v;
      else { free(rowbuf); return (*status = 410); }
    }
    done += n;
  }
  free(rowbuf);
  return 0;
}
