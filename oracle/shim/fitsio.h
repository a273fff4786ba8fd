/* TEST INFRASTRUCTURE ONLY -- stand-in for <fitsio.h> (cfitsio, unpinned).  Implements just enough of
 * the write path used by he_write_healpix_map (healpix_extra.c:132-164): an empty primary HDU plus one
 * BINTABLE extension with big-endian 1E columns in 2880-byte blocks; and of the read path of he_read_healpix_map
 * (healpix_extra.c:166-224): HDU walk, integer / string keys, rE / rD columns. */
#ifndef SHIM_FITSIO_H
#define SHIM_FITSIO_H
typedef struct shim_fitsfile fitsfile;
#define BINARY_TBL 2
#define TSTRING 16
#define TLONG 41
#define TFLOAT 42
#define TDOUBLE 82
#define READONLY 0
int fits_create_file(fitsfile **fptr, const char *filename, int *status);
int fits_create_tbl(fitsfile *fptr, int tbltype, long naxis2, int tfields, char **ttype, char **tform,
                    char **tunit, const char *extname, int *status);
int fits_write_key(fitsfile *fptr, int datatype, const char *keyname, void *value, const char *comm, int *status);
int fits_write_comment(fitsfile *fptr, const char *comm, int *status);
int fits_write_col(fitsfile *fptr, int datatype, int colnum, long firstrow, long firstelem, long nelem,
                   void *array, int *status);
int fits_close_file(fitsfile *fptr, int *status);
int fits_open_file(fitsfile **fptr, const char *filename, int mode, int *status);
int fits_movabs_hdu(fitsfile *fptr, int hdunum, int *exttype, int *status);
int fits_read_key_lng(fitsfile *fptr, const char *keyname, long *value, char *comm, int *status);
int fits_read_keys_lng(fitsfile *fptr, const char *keyname, int nstart, int nmax, long *value, int *nfound, int *status);
int fits_read_key(fitsfile *fptr, int datatype, const char *keyname, void *value, char *comm, int *status);
int fits_read_col(fitsfile *fptr, int datatype, int colnum, long firstrow, long firstelem, long nelem,
                  void *nulval, void *array, int *anynul, int *status);
#endif
