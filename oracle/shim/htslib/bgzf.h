/* TEST INFRASTRUCTURE ONLY.
 * Stand-in for htslib's <htslib/bgzf.h>, which is not installed in this image.
 * The reference's hot path uses exactly three BGZF symbols, all pure compression
 * (no arithmetic): bgzf_open / bgzf_write / bgzf_close (+ bgzf_mt in the exporter)
 * (reference src/importer/tailseq-import.c:55,68,83,110,282,294-297;
 *  src/importer/spotanalyzer.c:373,419-428; src/utils.c:56-76).
 * They are mapped onto zlib's gz* API; every consumer of these files reads them with
 * gzopen()/gzip.open(), so any valid gzip stream is equivalent for parity purposes. */
#ifndef TSK_ORACLE_BGZF_SHIM_H
#define TSK_ORACLE_BGZF_SHIM_H
#include <stdlib.h>
#include <sys/types.h>
#include <zlib.h>

typedef struct gzFile_s BGZF;

static inline BGZF *bgzf_open(const char *path, const char *mode)
{
    /* htslib's bgzf_open(path, "w") compresses at zlib's default level; keep that unless
     * TSK_SHIM_GZ_LEVEL=0..9 is set (bench.py sets it to 1 and says so, because
     * compression is outside the measured path). */
    const char *lvl = getenv("TSK_SHIM_GZ_LEVEL");
    char gzmode[8] = "wb";
    (void)mode;
    if (lvl != NULL && lvl[0] >= '0' && lvl[0] <= '9' && lvl[1] == '\0') {
        gzmode[2] = lvl[0];
        gzmode[3] = '\0';
    }
    return (BGZF *)gzopen(path, gzmode);
}

static inline ssize_t bgzf_write(BGZF *fp, const void *data, size_t length)
{
    int r;
    if (length == 0)
        return 0;
    r = gzwrite((gzFile)fp, data, (unsigned)length);
    return (r <= 0) ? -1 : (ssize_t)r;
}

/* tailseq-writefastq.c:387,399 asks for BGZF worker threads; the zlib stand-in has none */
static inline int bgzf_mt(BGZF *fp, int n_threads, int n_sub_blks)
{
    (void)fp; (void)n_threads; (void)n_sub_blks;
    return 0;
}

static inline int bgzf_close(BGZF *fp)
{
    return gzclose((gzFile)fp);
}
#endif
