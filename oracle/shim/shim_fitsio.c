/* TEST INFRASTRUCTURE ONLY -- minimal FITS BINTABLE writer behind the cfitsio names the reference's
 * he_write_healpix_map calls (healpix_extra.c:132-164).  Like cfitsio, fits_create_file fails when the
 * file already exists (status 105) and later calls are no-ops once status is non-zero. */
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
int fits_close_file(fitsfile *f, int *status)
{
  if (!f) return *status;
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

static int link_only(const char *w) { fprintf(stderr, "shim_fitsio: %s is link-only for GetHI\n", w); abort(); return 1; }
int fits_open_file(fitsfile **fptr, const char *filename, int mode, int *status)
{ (void)fptr; (void)filename; (void)mode; (void)status; return link_only("fits_open_file"); }
int fits_movabs_hdu(fitsfile *fptr, int hdunum, int *exttype, int *status)
{ (void)fptr; (void)hdunum; (void)exttype; (void)status; return link_only("fits_movabs_hdu"); }
int fits_read_key_lng(fitsfile *fptr, const char *keyname, long *value, char *comm, int *status)
{ (void)fptr; (void)keyname; (void)value; (void)comm; (void)status; return link_only("fits_read_key_lng"); }
int fits_read_keys_lng(fitsfile *fptr, const char *keyname, int nstart, int nmax, long *value, int *nfound, int *status)
{ (void)fptr; (void)keyname; (void)nstart; (void)nmax; (void)value; (void)nfound; (void)status; return link_only("fits_read_keys_lng"); }
int fits_read_key(fitsfile *fptr, int datatype, const char *keyname, void *value, char *comm, int *status)
{ (void)fptr; (void)datatype; (void)keyname; (void)value; (void)comm; (void)status; return link_only("fits_read_key"); }
int fits_read_col(fitsfile *fptr, int datatype, int colnum, long firstrow, long firstelem, long nelem, void *nulval,
                  void *array, int *anynul, int *status)
{ (void)fptr; (void)datatype; (void)colnum; (void)firstrow; (void)firstelem; (void)nelem; (void)nulval; (void)array; (void)anynul; (void)status; return link_only("fits_read_col"); }
