/* INI reader for the importer's per-tile configuration (the files scripts/generate-signalproc-conf.py
 * writes).  Written for this library as a span scanner over the whole file in memory; what it must
 * reproduce is the BEHAVIOUR the reference gets from its vendored parser as built by src/Makefile:3
 * (-DINI_MAX_LINE=1024, multi-line values, inline comments, BOM):
 *   - physical lines longer than 1023 bytes are seen as several lines (the reference reads with
 *     fgets into a 1024-byte buffer);
 *   - a line whose first non-blank character is ';' or '#' is a comment, indented or not;
 *   - an indented non-blank line after a "name = value" line continues that value: the handler is
 *     called again with the same name and the line's text (no inline-comment stripping);
 *   - "[section]": text up to the first ']' -- unless an inline comment (a ';' that follows a
 *     blank) comes first, in which case the line is ignored; section and remembered names are
 *     cut to 49 bytes;
 *   - "name = value": split at the first '=' (same inline-comment rule while looking for it), both
 *     sides trimmed, value cut at an inline comment;
 *   - the handler's return value is ignored on purpose: the reference's feeders return 0 on
 *     success, which its parser merely counts as an error without stopping, and parse_config()
 *     (parseconfig.c:695) only fails when the file cannot be opened -- unknown keys print a message.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

#define FGETS_PAYLOAD  1023      /* INI_MAX_LINE - 1 */
#define KEPT_NAME_LEN  49        /* MAX_SECTION - 1 == MAX_NAME - 1 */

struct span { const char *b, *e; };     /* [b, e) */

static int blank(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

static struct span trimmed(struct span s)
{
    while (s.e > s.b && blank(s.e[-1])) s.e--;
    while (s.b < s.e && blank(*s.b)) s.b++;
    return s;
}

/* where `stop` (0: nothing) first occurs in s, or where an inline comment starts, else s.e */
static const char *stop_or_comment(struct span s, char stop)
{
    const char *p;
    for (p = s.b; p < s.e; p++) {
        if (stop && *p == stop) return p;
        if (*p == ';' && p > s.b && blank(p[-1])) return p;
    }
    return s.e;
}

static void span_copy(char *dst, size_t cap, struct span s)
{
    size_t n = (size_t)(s.e - s.b);
    if (n > cap - 1) n = cap - 1;
    memcpy(dst, s.b, n);
    dst[n] = '\0';
}

int tsk_ini_parse(const char *filename, tsk_ini_handler handler, void *user)
{
    char section[KEPT_NAME_LEN + 1] = "", kept[KEPT_NAME_LEN + 1] = "";
    char name[FGETS_PAYLOAD + 1], value[FGETS_PAYLOAD + 1];
    char *text = NULL;
    size_t len = 0, cap = 0, got;
    const char *at, *end;
    int first = 1;
    FILE *fp = fopen(filename, "r");
    if (!fp) return -1;
    do {                                         /* the file is a few KB: take it whole */
        if (len + 4096 > cap) {
            char *q = realloc(text, cap = cap ? cap * 2 : 16384);
            if (!q) { free(text); fclose(fp); return -1; }
            text = q;
        }
        got = fread(text + len, 1, cap - len, fp);
        len += got;
    } while (got > 0);
    fclose(fp);

    for (at = text, end = text + len; at < end; first = 0) {
        /* one fgets() worth: through the newline, at most FGETS_PAYLOAD bytes; a NUL byte in the
         * file ends what the reference's string functions see of that line */
        struct span raw, ln;
        const char *nl = memchr(at, '\n', (size_t)(end - at)), *nul;
        raw.b = at;
        raw.e = nl ? nl + 1 : end;
        if (raw.e - raw.b > FGETS_PAYLOAD) raw.e = raw.b + FGETS_PAYLOAD;
        at = raw.e;
        if ((nul = memchr(raw.b, '\0', (size_t)(raw.e - raw.b))) != NULL) raw.e = nul;
        if (first && raw.e - raw.b >= 3 && !memcmp(raw.b, "\xEF\xBB\xBF", 3)) raw.b += 3;
        ln = trimmed(raw);
        if (ln.b == ln.e) continue;                             /* blank */
        if (*ln.b == ';' || *ln.b == '#') continue;             /* comment */
        if (kept[0] && ln.b > raw.b) {                          /* indented: more of the last value */
            span_copy(value, sizeof(value), ln);
            handler(user, section, kept, value);
        } else if (*ln.b == '[') {
            struct span in = {ln.b + 1, ln.e};
            const char *close = stop_or_comment(in, ']');
            if (close < in.e && *close == ']') {
                in.e = close;
                span_copy(section, sizeof(section), in);
                kept[0] = '\0';
            }
        } else {
            const char *eq = stop_or_comment(ln, '=');
            if (eq < ln.e && *eq == '=') {
                struct span key = {ln.b, eq}, val = {eq + 1, ln.e};
                while (key.e > key.b && blank(key.e[-1])) key.e--;
                while (val.b < val.e && blank(*val.b)) val.b++;
                val.e = stop_or_comment(val, 0);
                while (val.e > val.b && blank(val.e[-1])) val.e--;
                span_copy(name, sizeof(name), key);
                span_copy(value, sizeof(value), val);
                span_copy(kept, sizeof(kept), key);
                handler(user, section, name, value);
            }
        }
    }
    free(text);
    return 0;
}
