/* TEST INFRASTRUCTURE ONLY -- prints what the reference's vendored INI parser (src/contrib/ini.c,
 * compiled where it lies with the reference's -DINI_MAX_LINE=1024) hands to a handler, in the
 * format of tailseeker_b200/bin/tsk-ini-dump.  The handler returns 0 like the reference's feeders. */
#include <stdio.h>
#include "ini.h"

static int show(void *user, const char *section, const char *name, const char *value)
{
    (void)user;
    printf("%s\x1f%s\x1f%s\n", section, name, value);
    return 0;
}

int main(int argc, char *argv[])
{
    if (argc != 2) return 2;
    return ini_parse(argv[1], show, NULL) < 0 ? 1 : 0;
}
