/* Minimal INI reader with the conventions the reference's configuration relies on
 * (src/contrib/ini.c as configured by src/Makefile:3): "[section]" lines, "name = value" pairs,
 * whole-line comments starting with ';' or '#', inline comments introduced by " ;", lines up to
 * 1024 bytes.  The handler's return value is ignored on purpose: the reference's feeders return 0
 * on success, which its parser counts as an error without stopping (parseconfig.c:695 only
 * aborts when the file cannot be opened), so unknown keys merely print a message. */
#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include "tsk_host.h"

#define LINE_MAX_LEN 1024

static char *rstrip(char *s)
{
    char *p = s + strlen(s);
    while (p > s && isspace((unsigned char)p[-1])) *--p = '\0';
    return s;
}

static char *lskip(char *s)
{
    while (*s && isspace((unsigned char)*s)) s++;
    return s;
}

/* first occurrence of one of `chars`, or of an inline comment (" ;"), else the terminator */
static char *find_chars_or_comment(char *s, const char *chars)
{
    int was_space = 0;
    while (*s && (!chars || !strchr(chars, *s)) && !(was_space && *s == ';')) {
        was_space = isspace((unsigned char)*s);
        s++;
    }
    return s;
}

int tsk_ini_parse(const char *filename, tsk_ini_handler handler, void *user)
{
    char line[LINE_MAX_LEN], section[64] = "", prev_name[64] = "";
    FILE *fp = fopen(filename, "r");
    if (!fp) return -1;
    while (fgets(line, sizeof(line), fp)) {
        char *start = lskip(rstrip(line)), *end;
        if (*start == ';' || *start == '#' || *start == '\0')
            continue;
        if (*prev_name && start > line) {          /* continuation of the previous value */
            handler(user, section, prev_name, start);
        } else if (*start == '[') {
            end = find_chars_or_comment(start + 1, "]");
            if (*end == ']') {
                *end = '\0';
                strncpy(section, start + 1, sizeof(section) - 1);
                section[sizeof(section) - 1] = '\0';
                *prev_name = '\0';
            }
        } else {
            end = find_chars_or_comment(start, "=");
            if (*end == '=') {
                char *name, *value;
                *end = '\0';
                name = rstrip(start);
                value = lskip(end + 1);
                end = find_chars_or_comment(value, NULL);
                if (*end) *end = '\0';
                rstrip(value);
                strncpy(prev_name, name, sizeof(prev_name) - 1);
                prev_name[sizeof(prev_name) - 1] = '\0';
                handler(user, section, name, value);
            }
        }
    }
    fclose(fp);
    return 0;
}
