/*
 * sw_io.c -- pure-host helpers of libsw_b200.so: 2-bit packing, FASTA input and the two
 * text layouts the reference's golden files use.  No scoring happens here.
 *
 * Mirrors: aligner_Header.c:14-47 (charTo2bit), ScoreBank_v1_tb.sv:184-216 (FASTA tokens),
 * ScoreBank_v1_tb.sv:280-281 (result line), data/score500.txt:1-3 (ssearch36 -R rows).
 */
#include "../../include/sw_b200.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

static unsigned code_of(char c)
{
    switch (c) {                       /* aligner_Header.c:34-39 */
        case 'a': case 'A': return 2u;
        case 'c': case 'C': return 1u;
        case 'g': case 'G': return 3u;
        default: return 0u;            /* T and every unknown letter */
    }
}

void sw_pack_2bit(const char *seq, size_t len, uint8_t *out)
{
    size_t i;
    memset(out, 0, (len + 3) / 4);
    for (i = 0; i < len; ++i) out[i >> 2] |= (uint8_t)(code_of(seq[i]) << ((i * 2) & 7));
}

void sw_unpack_2bit(const uint8_t *packed, size_t len, char *out)
{
    static const char letters[4] = {'T', 'C', 'A', 'G'};
    size_t i;
    for (i = 0; i < len; ++i) out[i] = letters[(packed[i >> 2] >> ((i * 2) & 7)) & 3];
    out[len] = '\0';
}

void sw_seqset_free(sw_seqset_t *s)
{
    size_t i;
    if (!s) return;
    if (s->name) for (i = 0; i < s->n; ++i) free(s->name[i]);
    free(s->name); free(s->packed); free(s->len); free(s->off); free(s);
}

/* growable record list */
typedef struct { char *name; char *seq; size_t len, cap; } rec_t;

static int rec_append(rec_t *r, const char *s, size_t n)
{
    if (r->len + n + 1 > r->cap) {
        size_t nc = (r->cap ? r->cap * 2 : 256);
        char *p;
        while (nc < r->len + n + 1) nc *= 2;
        p = (char *)realloc(r->seq, nc);
        if (!p) return -1;
        r->seq = p; r->cap = nc;
    }
    memcpy(r->seq + r->len, s, n);
    r->len += n;
    r->seq[r->len] = '\0';
    return 0;
}

int sw_read_fasta(const char *path, sw_seqset_t **out)
{
    FILE *f;
    char *line = NULL;
    size_t lcap = 0, nrec = 0, rcap = 0, i, total = 0;
    rec_t *recs = NULL;
    int rc = SW_OK, headerless = 0;
    sw_seqset_t *s = NULL;

    if (!path || !out) return SW_EINVAL;
    *out = NULL;
    f = fopen(path, "r");
    if (!f) return SW_EIO;
    {
        /* read whole lines of any length */
        int ch;
        size_t ll = 0;
        for (;;) {
            ch = fgetc(f);
            if (ch != EOF && ch != '\n') {
                if (ll + 2 > lcap) {
                    size_t nc = lcap ? lcap * 2 : 512;
                    char *p = (char *)realloc(line, nc);
                    if (!p) { rc = SW_ENOMEM; break; }
                    line = p; lcap = nc;
                }
                line[ll++] = (char)ch;
                continue;
            }
            if (line) line[ll] = '\0';
            if (ll > 0) {
                /* trim */
                char *b = line;
                size_t n = ll;
                while (n && isspace((unsigned char)b[n - 1])) b[--n] = '\0';
                while (*b && isspace((unsigned char)*b)) { ++b; --n; }
                if (n > 0) {
                    if (b[0] == '>') {
                        char *e = b + 1;
                        if (nrec == rcap) {
                            size_t nc = rcap ? rcap * 2 : 64;
                            rec_t *p = (rec_t *)realloc(recs, nc * sizeof(rec_t));
                            if (!p) { rc = SW_ENOMEM; break; }
                            recs = p; rcap = nc;
                        }
                        while (*e && !isspace((unsigned char)*e)) ++e;   /* name = first token */
                        *e = '\0';
                        memset(&recs[nrec], 0, sizeof(rec_t));
                        recs[nrec].name = strdup(b + 1);
                        if (!recs[nrec].name) { rc = SW_ENOMEM; break; }
                        ++nrec;
                    } else {
                        if (nrec == 0) {
                            /* no header: main_test.c:304 reads the first token as the sequence */
                            recs = (rec_t *)calloc(1, sizeof(rec_t));
                            if (!recs) { rc = SW_ENOMEM; break; }
                            rcap = 1; nrec = 1; headerless = 1;
                            recs[0].name = strdup("seq0");
                        }
                        if (headerless && recs[0].len > 0) {
                            /* only the first token counts in header-less files */
                        } else {
                            /* drop inner whitespace, keep letters */
                            char *w = b, *r = b;
                            for (; *r; ++r) if (!isspace((unsigned char)*r)) *w++ = *r; else if (headerless) break;
                            *w = '\0';
                            if (rec_append(&recs[nrec - 1], b, (size_t)(w - b)) != 0) { rc = SW_ENOMEM; break; }
                        }
                    }
                }
            }
            ll = 0;
            if (ch == EOF) break;
        }
    }
    fclose(f);
    free(line);
    if (rc == SW_OK) {
        s = (sw_seqset_t *)calloc(1, sizeof(sw_seqset_t));
        if (!s) rc = SW_ENOMEM;
    }
    if (rc == SW_OK) {
        for (i = 0; i < nrec; ++i) total += (recs[i].len + 3) / 4;
        s->n = nrec;
        s->packed = (uint8_t *)calloc(total + 16, 1);
        s->len = (uint32_t *)calloc(nrec ? nrec : 1, sizeof(uint32_t));
        s->off = (uint64_t *)calloc(nrec ? nrec : 1, sizeof(uint64_t));
        s->name = (char **)calloc(nrec ? nrec : 1, sizeof(char *));
        s->packed_bytes = total;
        if (!s->packed || !s->len || !s->off || !s->name) rc = SW_ENOMEM;
    }
    if (rc == SW_OK) {
        size_t o = 0;
        for (i = 0; i < nrec; ++i) {
            s->len[i] = (uint32_t)recs[i].len;
            s->off[i] = o;
            if (recs[i].len) sw_pack_2bit(recs[i].seq, recs[i].len, s->packed + o);
            o += (recs[i].len + 3) / 4;
            s->name[i] = recs[i].name;
            recs[i].name = NULL;
        }
        *out = s;
        s = NULL;
    }
    for (i = 0; i < nrec; ++i) { free(recs[i].name); free(recs[i].seq); }
    free(recs);
    if (s) sw_seqset_free(s);
    return rc;
}

int sw_write_out_txt(FILE *f, const sw_seqset_t *db, const int32_t *scores, const uint64_t *time_ns)
{
    size_t i;
    if (!f || !db || !scores) return SW_EINVAL;
    for (i = 0; i < db->n; ++i) {
        char nm[512];
        snprintf(nm, sizeof nm, ">%s", db->name[i] ? db->name[i] : "");
        /* Verilog "%10s" right-aligns; "%d" of a 32-bit integer prints 11 columns */
        if (fprintf(f, "@%6lluns: %10s score: \t%11d\n",
                    (unsigned long long)(time_ns ? time_ns[i] : 0ull), nm, (int)scores[i]) < 0)
            return SW_EIO;
    }
    return SW_OK;
}

int sw_write_ssearch_R(FILE *f, const char *query_file, const char *db_file, const sw_seqset_t *query,
                       const sw_seqset_t *db, const int32_t *scores)
{
    size_t i;
    unsigned long long offs = 0;
    if (!f || !query || !db || !scores || query->n < 1) return SW_EINVAL;
    fprintf(f, "# sw_b200 -3 -R -n %s %s\n", query_file ? query_file : "-", db_file ? db_file : "-");
    fprintf(f, ">>>0 %u\t%s - %u nt\n", query->len[0], query->name[0] ? query->name[0] : "query", query->len[0]);
    for (i = 0; i < db->n; ++i) {
        /* name, length, frame, two unused statistics, SCORE (6th field), then the fields
           ssearch36 uses for its own bookkeeping (kept so that column positions match) */
        if (fprintf(f, "%-12s %6u 0 -1.00000 -1.00000 %4d    0    0  1  0    0    0    0  1  0 %5llu %8llu\n",
                    db->name[i] ? db->name[i] : "", db->len[i], (int)scores[i],
                    (unsigned long long)i, offs) < 0)
            return SW_EIO;
        offs += strlen(db->name[i] ? db->name[i] : "") + 2 + db->len[i] + 1;
    }
    fprintf(f, "# %llu sequences\n", (unsigned long long)db->n);
    return SW_OK;
}
