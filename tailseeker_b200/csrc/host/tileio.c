/*
 * tileio.c -- Illumina per-cycle file readers, kept in the files' native layouts:
 *   CIF  "CIF", u8 version=1, u8 datasize=2, u16 first_cycle, u16 ncycles, u32 N, then four planes
 *        int16[N] (A,C,G,T)              (reference src/importer/cifreader.c:38-102,118-170,235)
 *   BCL  u32 N, u8[N]; ".bcl" or ".bcl.gz" (reference src/importer/bclreader.c:43-80,95-119,206)
 * Blocks are read straight into bcl[T][count] / cif[L3][4][count]: no AoS transpose.
 * Colour matrix: 16 numbers, inverted like src/contrib/misc.c:11-58 (float cofactors, same order).
 */
#include <endian.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

struct tile_files {
    int T, L3;
    uint32_t nclusters, read;
    FILE **cif;
    gzFile *bcl;
};

static float cofactor_entry(int i, int j, const float *m)
{
    static const signed char term[6][6] = {
        {+1, -1, 0, 0, -1, +1}, {+1, +1, 0, -1, -1, 0}, {-1, -1, +1, 0, 0, +1},
        {-1, -1, 0, 0, +1, +1}, {-1, +1, 0, -1, +1, 0}, {+1, -1, -1, 0, 0, +1},
    };
    const int o = 2 + (j - i), ci = i + 4 + o, rj = j + 4 - o;
    float acc = 0.f;
    for (int t = 0; t < 6; t++) {
        const float prod = m[((rj + term[t][1]) % 4) * 4 + ((ci + term[t][0]) % 4)] *
                           m[((rj + term[t][3]) % 4) * 4 + ((ci + term[t][2]) % 4)] *
                           m[((rj + term[t][5]) % 4) * 4 + ((ci + term[t][4]) % 4)];
        if (t == 0) acc = prod;
        else if (t < 3) acc = acc + prod;
        else acc = acc - prod;
    }
    return (o % 2) ? acc : -acc;
}

int tsk_invert_matrix(const float m[16], float out[16])
{
    float adj[16], det = 0.f;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            adj[j * 4 + i] = cofactor_entry(i, j, m);
    for (int k = 0; k < 4; k++) det += m[k] * adj[k * 4];
    if (det == 0) return -1;
    det = 1. / det;
    for (int i = 0; i < 16; i++) out[i] = adj[i] * det;
    return 0;
}

int tsk_load_color_matrix(float inv[16], const char *filename)
{
    float m[16];
    FILE *fp = fopen(filename, "rt");
    if (!fp) { fprintf(stderr, "Color matrix <%s> could not be opened.\n", filename); return -1; }
    for (int i = 0; i < 16; i++)
        if (fscanf(fp, "%f", &m[i]) < 1) {
            fclose(fp);
            fprintf(stderr, "Color matrix <%s> could not be loaded.\n", filename);
            return -1;
        }
    fclose(fp);
    if (tsk_invert_matrix(m, inv) < 0) { fprintf(stderr, "Color matrix could not be inverted.\n"); return -1; }
    return 0;
}

static FILE *open_cif(const char *path, uint32_t *nclusters)
{
    unsigned char h[13];
    FILE *fp = fopen(path, "rb");
    if (!fp) { fprintf(stderr, "open_cif_file: Can't open file %s.\n", path); return NULL; }
    if (fread(h, 1, 13, fp) != 13) { fprintf(stderr, "Unexpected EOF %s.\n", path); fclose(fp); return NULL; }
    if (memcmp(h, "CIF", 3) != 0) { fprintf(stderr, "File %s does not seem to be a CIF file.", path); fclose(fp); return NULL; }
    if (h[3] != 1) { fprintf(stderr, "Unsupported CIF version: %d.\n", h[3]); fclose(fp); return NULL; }
    if (h[4] != 2) { fprintf(stderr, "Unsupported data size (%d).\n", h[4]); fclose(fp); return NULL; }
    *nclusters = (uint32_t)h[9] | ((uint32_t)h[10] << 8) | ((uint32_t)h[11] << 16) | ((uint32_t)h[12] << 24);
    return fp;
}

struct tile_files *tsk_open_tile(const struct host_config *cfg, const char *msgprefix)
{
    char path[PATH_MAX];
    struct tile_files *tf = calloc(1, sizeof(*tf));
    if (!tf) return NULL;
    tf->T = cfg->total_cycles; tf->L3 = cfg->threep_length;
    tf->cif = calloc(tf->L3, sizeof(FILE *));
    tf->bcl = calloc(tf->T, sizeof(gzFile));
    printf("%sUsing cluster intensities for cycle %d-%d from %s.\n", msgprefix, cfg->threep_start + 1,
           cfg->threep_start + cfg->threep_length, cfg->datadir);
    for (int i = 0; i < tf->L3; i++) {
        uint32_t n = 0;
        snprintf(path, PATH_MAX, "%s/L%03d/C%d.1/s_%d_%04d.cif", cfg->datadir, cfg->lane,
                 cfg->threep_start + i + 1, cfg->lane, cfg->tile);
        tf->cif[i] = open_cif(path, &n);
        if (!tf->cif[i]) { tsk_close_tile(tf); return NULL; }
        if (i == 0) tf->nclusters = n;
        else if (n != tf->nclusters) {
            fprintf(stderr, "Inconsistent number of clusters in CIF cycle %d.\n", cfg->threep_length + i + 1);
            tsk_close_tile(tf);
            return NULL;
        }
    }
    for (int c = 0; c < tf->T; c++) {
        uint32_t n;
        snprintf(path, PATH_MAX, "%s/BaseCalls/L%03d/C%d.1/s_%d_%04d.bcl", cfg->datadir, cfg->lane, c + 1,
                 cfg->lane, cfg->tile);
        tf->bcl[c] = gzopen(path, "rb");
        if (!tf->bcl[c]) {
            strncat(path, ".gz", PATH_MAX - strlen(path) - 1);
            tf->bcl[c] = gzopen(path, "rb");
        }
        if (!tf->bcl[c]) { fprintf(stderr, "load_bcl_file: Can't open file %s.\n", path); tsk_close_tile(tf); return NULL; }
        if (gzread(tf->bcl[c], &n, 4) < 4) { fprintf(stderr, "Unexpected EOF %s.\n", path); tsk_close_tile(tf); return NULL; }
        n = le32toh(n);
        if (n != tf->nclusters) {
            fprintf(stderr, "Inconsistent number of clusters in cycle %d.\n", c + 1);
            tsk_close_tile(tf);
            return NULL;
        }
    }
    printf("%sUsing BCL base calls for cycle %d-%d.\n", msgprefix, 1, tf->T);
    return tf;
}

uint32_t tsk_tile_nclusters(const struct tile_files *tf) { return tf->nclusters; }

int tsk_read_block(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif)
{
    for (int i = 0; i < tf->L3; i++)
        for (int ch = 0; ch < 4; ch++) {
            const long at = 13 + ((long)tf->nclusters * ch + tf->read) * 2;
            int16_t *dst = cif + ((size_t)i * 4 + ch) * count;
            if (fseek(tf->cif[i], at, SEEK_SET) != 0) {
                fprintf(stderr, "Failed to seek cluster intensity values in a CIF.\n");
                return -1;
            }
            if (fread(dst, 2, count, tf->cif[i]) != count) {
                fprintf(stderr, "Not all data were loaded from a CIF.\n");
                return -1;
            }
#if __BYTE_ORDER != __LITTLE_ENDIAN
            for (uint32_t k = 0; k < count; k++) dst[k] = (int16_t)le16toh((uint16_t)dst[k]);
#endif
        }
    for (int c = 0; c < tf->T; c++) {
        uint8_t *dst = bcl + (size_t)c * count;
        uint32_t got = 0;
        while (got < count) {          /* gzread takes an unsigned int length */
            const uint32_t want = count - got > (1u << 30) ? (1u << 30) : count - got;
            const int r = gzread(tf->bcl[c], dst + got, want);
            if (r < 1) { fprintf(stderr, "Unexpected EOF in a BCL (cycle %d).\n", c + 1); return -1; }
            got += (uint32_t)r;
        }
    }
    tf->read += count;
    return 0;
}

void tsk_close_tile(struct tile_files *tf)
{
    if (!tf) return;
    if (tf->cif) for (int i = 0; i < tf->L3; i++) if (tf->cif[i]) fclose(tf->cif[i]);
    if (tf->bcl) for (int c = 0; c < tf->T; c++) if (tf->bcl[c]) gzclose(tf->bcl[c]);
    free(tf->cif); free(tf->bcl); free(tf);
}
