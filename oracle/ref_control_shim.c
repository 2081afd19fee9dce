/* TEST INFRASTRUCTURE ONLY.  A three-function doorway into the UNMODIFIED reference control
 * aligner, so that tests can ask it about single reads.  Compiled by oracle/Makefile together
 * with the reference's own controlaligner.c, phix_control.c, ssw.c and my_strstr.c (where they
 * lie under /root/reference) into oracle/_ref/libtskref_control.so; nothing here restates any
 * algorithm. */
#include <string.h>
#include "tailseq-import.h"
#include "../contrib/ssw.h"

static struct ControlFilterInfo ctl;
static int ctl_ready = 0;

/* initialize_control_aligner(), controlaligner.c:148-171 */
int tskref_control_setup(int first_cycle, int read_length)
{
    if (ctl_ready) free_control_aligner(&ctl);
    memset(&ctl, 0, sizeof(ctl));
    strcpy(ctl.name, "PhiX");
    ctl.first_cycle = first_cycle;
    ctl.read_length = read_length;
    ctl_ready = initialize_control_aligner(&ctl) == 0;
    return ctl_ready ? ctl.min_control_alignment_score : -1;
}

/* try_alignment_to_control(), controlaligner.c:107-145 */
int tskref_control_try(const char *sequence_read)
{
    return try_alignment_to_control(&ctl, sequence_read);
}

/* score1 of the ssw_align() call at controlaligner.c:127-133, without the fast path */
int tskref_control_score(const char *sequence_read)
{
    static const int8_t code[128] = {
        ['A'] = 0, ['C'] = 1, ['G'] = 2, ['T'] = 3, ['N'] = 4,
    };
    int8_t read_seq[256];
    s_profile *prof;
    s_align *aln;
    int score = -1;
    for (int i = 0; i < ctl.read_length; i++)
        read_seq[i] = code[(int)sequence_read[ctl.first_cycle + i]];
    prof = ssw_init(read_seq, ctl.read_length, ctl.ssw_score_mat, 5, 0);
    aln = ssw_align(prof, ctl.control_seq, ctl.control_seq_length, CONTROL_ALIGN_GAP_OPEN_SCORE,
                    CONTROL_ALIGN_GAP_EXTENSION_SCORE, 2, ctl.min_control_alignment_score, 0,
                    ctl.control_alignment_mask_len);
    if (aln != NULL) { score = aln->score1; align_destroy(aln); }
    init_destroy(prof);
    return score;
}
