/* tsk-ini-dump {file.ini} -- prints every (section, name, value) the importer's INI reader hands to
 * the configuration feeder, one per line, fields separated by 0x1f.  A maintainer's view of ini.c;
 * tests/test_ini.py compares it with the reference's vendored parser on awkward files. */
#include <stdio.h>
#include "tsk_host.h"

static int show(void *user, const char *section, const char *name, const char *value)
{
    (void)user;
    printf("%s\x1f%s\x1f%s\n", section, name, value);
    return 0;
}

int main(int argc, char *argv[])
{
    if (argc != 2) { fprintf(stderr, "Usage: %s {file.ini}\n", argv[0]); return 2; }
    if (tsk_ini_parse(argv[1], show, NULL) < 0) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    return 0;
}
