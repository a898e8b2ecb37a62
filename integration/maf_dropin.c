/* maf_dropin.c -- the reference-side binding of MAF block input and output: multiz's own `mafNext`,
 * `mafReadAll` and `mafWrite` symbols (maf.h:64-79), re-implemented around whole-buffer I/O.  The reference's definitions
 * (maf.c:133-216, :219-230, :251-294) stay in maf.o under other names (maf.c is compiled with -DmafNext=ref_mafNext
 * -DmafReadAll=ref_mafReadAll -DmafWrite=ref_mafWrite, integration/Makefile); YB_MAF=ref routes back to them.
 *
 * What the reference does per block: a line arrives through fgetc() one character at a time with a buffer-size check per
 * character (get_line, maf.c:53-70), an `s` line is taken apart by sscanf("s %s %d %d %c %d %s") -- which walks the line
 * twice more -- followed by strlen and a dash count; a block leaves through seven fprintf calls per row.  Once the
 * alignment itself takes milliseconds, that is the wall clock of a merge.
 *
 * Here: the file is read in 1 MiB pieces, a line is found with memchr, and ONE pass over an `s` line's text finds its end
 * (the first white-space byte, what %s stops at) and counts its dashes.  Fields are parsed by hand only when that is
 * provably what sscanf would return -- unsigned decimal numbers of at most nine digits separated by blanks or tabs;
 * anything else (signs, overflow, other white space, missing fields) goes through the reference's own sscanf call, so
 * values and failures are the reference's by construction.  Every message, check, allocation and counter update of
 * maf.c:133-216 is kept in its order.  A block is formatted into one buffer -- same widths, same name re-joining
 * (parseSrcName, multi_util.c:889-906), the score through the reference's own "%3.1f" -- and written with one fwrite.
 *
 * C, because it fills the reference's own structs (maf.h:13-58), included from the reference tree.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "maf.h"

void fatalf(const char *fmt, ...);                    /* util.c:21 */
void fatal(const char *msg);                          /* util.c:17 */
void *ckalloc(size_t amount);                         /* util.c:41 */
char *copy_string(const char *s);                     /* util.c:81 */
void parseSrcName2(struct mafComp *c);                /* multi_util.c:909 */
int parseScoreLine(char *line, struct mafAli *ali);   /* maf.c:88 */
int digitsBaseTen(int x);                             /* maf.c:239 (fatal on a negative number) */
struct mafAli *ref_mafNext(struct mafFile *mf);       /* the reference's, renamed */
void ref_mafWrite(FILE *f, struct mafAli *a);
int yb_host_fclose(FILE *f);                          /* yama_dropin.cpp: what the reference objects' fclose() is */

static int use_ref(void) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("YB_MAF"); v = (e && strcmp(e, "ref") == 0) ? 1 : 0; }
    return v;
}

/* ---- line source: the rest of mf->fp (mafOpen has taken the header line with fgets, maf.c:18), in large pieces --------- */
struct reader {
    struct reader *next;
    struct mafFile *mf;
    FILE *fp;               /* the stream the buffered bytes came from: a mafFile freed before its end and another one
                               allocated at the same address must not inherit them */
    char *buf;              /* cap + 1 bytes: a line is NUL-terminated in place */
    size_t cap, lo, hi;     /* unread bytes: buf[lo, hi) */
    char saved;             /* the byte the last line's terminator replaced */
    int have_saved, eof;
};
static struct reader *readers = NULL;

static struct reader *reader_of(struct mafFile *mf) {
    struct reader *r;
    for (r = readers; r != NULL; r = r->next)
        if (r->mf == mf) {
            if (r->fp == mf->fp) return r;
            r->fp = mf->fp; r->lo = r->hi = 0; r->have_saved = 0; r->eof = 0;      /* a new file behind an old address */
            return r;
        }
    r = ckalloc(sizeof *r);
    r->mf = mf; r->fp = mf->fp; r->cap = (size_t)1 << 20; r->buf = ckalloc(r->cap + 1);
    r->lo = r->hi = 0; r->have_saved = 0; r->eof = 0; r->saved = 0;
    r->next = readers; readers = r;
    return r;
}
static void reader_drop(struct reader *r) {
    struct reader **pp;
    for (pp = &readers; *pp != NULL; pp = &(*pp)->next)
        if (*pp == r) { *pp = r->next; break; }
    free(r->buf);
    free(r);
}

/* get_line (maf.c:53-70): the next line with its '\n' (the last one may lack it), NUL-terminated; its length, or -1 at
 * the end of the file */
static long next_line(struct reader *r, FILE *fp, char **linep) {
    char *nl;
    size_t n;
    if (r->have_saved) { r->buf[r->lo] = r->saved; r->have_saved = 0; }
    for (;;) {
        nl = r->hi > r->lo ? memchr(r->buf + r->lo, '\n', r->hi - r->lo) : NULL;
        if (nl != NULL || r->eof) break;
        if (r->lo > 0) {                                     /* keep the partial line, refill behind it */
            memmove(r->buf, r->buf + r->lo, r->hi - r->lo);
            r->hi -= r->lo; r->lo = 0;
        }
        if (r->hi == r->cap) {
            r->cap *= 2;
            if ((r->buf = realloc(r->buf, r->cap + 1)) == NULL) return -1;      /* (maf.c:44-46: treated as the end) */
        }
        n = fread(r->buf + r->hi, 1, r->cap - r->hi, fp);
        if (n == 0) r->eof = 1;
        r->hi += n;
    }
    if (nl == NULL && r->hi == r->lo) return -1;
    n = nl != NULL ? (size_t)(nl - (r->buf + r->lo)) + 1 : r->hi - r->lo;
    *linep = r->buf + r->lo;
    r->lo += n;
    r->saved = r->buf[r->lo]; r->have_saved = 1;             /* (buf has cap + 1 bytes) */
    r->buf[r->lo] = 0;
    return (long)n;
}

/* get_maf_line (maf.c:74-88): comment lines are counted, echoed (verbose, unless they mention "eof") and skipped */
static long next_maf_line(struct reader *r, FILE *fp, struct mafFile *mf, char **linep) {
    long nn;
    while ((nn = next_line(r, fp, linep)) > 1) {
        mf->line_nbr++;
        if ((*linep)[0] == '#') {
            if (mf->verbose && strstr(*linep, "eof") == NULL)
                printf("%s", *linep);
        } else
            break;
    }
    return nn;
}

/* bit 0: a dash; bit 1: where %s stops (white space as isspace() in the C locale sees it, or the end of the string) */
static unsigned char cls[256];
static void cls_init(void) {
    if (cls[0]) return;
    cls['-'] = 1;
    cls[' '] = cls['\t'] = cls['\n'] = cls['\v'] = cls['\f'] = cls['\r'] = 2;
    cls[0] = 2;
}

static const char *skip_blank(const char *p) {
    while (*p == ' ' || *p == '\t') ++p;
    return p;
}
/* an unsigned decimal of 1..9 digits followed by a blank or tab: what %d returns for it, without its corner cases */
static int plain_number(const char **pp, int *out) {
    const char *p = *pp;
    int v = 0, nd = 0;
    while (*p >= '0' && *p <= '9' && nd < 10) { v = v * 10 + (*p - '0'); ++p; ++nd; }
    if (nd < 1 || nd > 9 || (*p != ' ' && *p != '\t')) return 0;
    *out = v; *pp = p;
    return 1;
}

/* One `s` line (maf.c:168-172).  Returns 1 with c->start/size/strand/srcSize, the source name in buf and the text in
 * c->text (its length in *tlen, its non-dash count in *nondash) when the line is plain; 0 when sscanf has to decide. */
static int plain_component(const char *line, char *buf, struct mafComp *c, size_t *tlen, int *nondash) {
    const char *p = skip_blank(line + 1), *q;
    const unsigned char *t;
    size_t n, dashes = 0;
    for (q = p; !(cls[(unsigned char)*q] & 2); ++q) {}
    n = (size_t)(q - p);
    if (n < 1 || n > 499 || (*q != ' ' && *q != '\t')) return 0;
    memcpy(buf, p, n); buf[n] = 0;
    p = skip_blank(q);
    if (!plain_number(&p, &c->start)) return 0;
    p = skip_blank(p);
    if (!plain_number(&p, &c->size)) return 0;
    p = skip_blank(p);
    if (cls[(unsigned char)*p] & 2) return 0;
    c->strand = *p++;
    if (*p != ' ' && *p != '\t') return 0;
    p = skip_blank(p);
    if (!plain_number(&p, &c->srcSize)) return 0;
    p = skip_blank(p);
    t = (const unsigned char *)p;
    while (!(cls[*t] & 2)) { dashes += cls[*t]; ++t; }
    n = (size_t)(t - (const unsigned char *)p);
    if (n < 1) return 0;
    memcpy(c->text, p, n); c->text[n] = 0;
    *tlen = n; *nondash = (int)(n - dashes);
    return 1;
}

struct mafAli *mafNext(struct mafFile *mf) {
    FILE *fp;
    struct reader *r;
    struct mafAli *a;
    struct mafComp *c, *last;
    char buf[500], blockHeaderLine[1000];
    char *line = NULL;
    int i, len;

    if (use_ref()) return ref_mafNext(mf);
    cls_init();
    fp = mf->fp;
    a = ckalloc(sizeof(struct mafAli));
    r = reader_of(mf);
    while ((len = next_maf_line(r, fp, mf, &line)) != -1)
        if (line[0] != '#' && line[0] != '\n' && line[0] != ' ')
            break;
    if (len == -1) {
        reader_drop(r);
        yb_host_fclose(fp);
        mf->fp = NULL;
        return NULL;
    }
    if (strncmp(line, "a", 1) == 0)
        strcpy(blockHeaderLine, line);
    else
        fatalf("Expecting 'a (score=xxx)' in file %s, line %d:\n%s",
               mf->fileName, mf->line_nbr, line);
    a->textSize = 0;
    last = a->components = NULL;
    a->next = NULL;
    while ((len = next_maf_line(r, fp, mf, &line)) != -1 &&
            line[0] != '\n' && line[0] != ' ' && line[0] != '#') {
        size_t tlen = 0;
        int nondash = -1;
        c = ckalloc(sizeof(struct mafComp));
        c->text = ckalloc(len * sizeof(char));
        if (line[0] != 's')
            continue;
        if (!plain_component(line, buf, c, &tlen, &nondash)) {
            nondash = -1;
            if (sscanf(line, "s %s %d %d %c %d %s",
                       buf, &(c->start), &(c->size), &(c->strand),
                       &(c->srcSize), c->text) != 6)
                fatalf("bad component in file %s, line %d:\n%s",
                       mf->fileName, mf->line_nbr, buf);
            tlen = strlen(c->text);
        }
        c->src = copy_string(buf);
        parseSrcName2(c);
        c->paralog = 's';
        c->mafPosMap = NULL;
        c->next = NULL;
        if (a->components == NULL) {
            a->textSize = (int)tlen;
            a->components = c;
        } else {
            if (a->textSize != (signed)tlen)
                fatalf("line %d of %s: inconsistent row size",
                       mf->line_nbr, mf->fileName);
            last->next = c;
        }
        last = c;
        if (c->srcSize <= 0 || c->size <= 0)
            fatalf("Size <= 0 at line %d of file %s:\n%s",
                   mf->line_nbr, mf->fileName, line);
        if (c->start < 0 || c->start + c->size > c->srcSize) {
            if (c != a->components) {
                c = a->components;
                fprintf(stderr,
                        "in maf entry with top row %s:%d len = %d,\n",
                        c->src, c->start, c->size);
            }
            fatalf("Bad coordinates at line %d of file %s:\n%s",
                   mf->line_nbr, mf->fileName, line);
        }
        if (nondash >= 0)
            len = nondash;
        else
            for (i = len = 0; i < a->textSize; ++i)
                if (c->text[i] != '-')
                    ++len;
        if (len != c->size)
            fatalf("Actual size %d, claimed size %d at line %d of file %s:\n%s", len, c->size, mf->line_nbr, mf->fileName, line);
    }
    parseScoreLine(blockHeaderLine, a);
    mf->line_nbr++;
    return a;
}

struct mafFile *mafReadAll(char *fileName, int verbose) {          /* maf.c:219-230 */
    struct mafFile *mf = mafOpen(fileName, verbose);
    struct mafAli *a, *last;

    for (last = NULL; (a = mafNext(mf)) != NULL; last = a)
        if (last == NULL)
            mf->alignments = a;
        else
            last->next = a;
    return mf;
}

/* ---- output (maf.c:251-294) --------------------------------------------------------------------------------------------- */
static char *wbuf = NULL;
static size_t wcap = 0, wlen = 0;
static void wneed(size_t more) {
    if (wlen + more <= wcap) return;
    while (wlen + more > wcap) wcap = wcap ? 2 * wcap : (size_t)1 << 16;
    if ((wbuf = realloc(wbuf, wcap)) == NULL) fatal("mafWrite: out of memory");
}
static void wpad(size_t n) { memset(wbuf + wlen, ' ', n); wlen += n; }
/* "%*d" of a non-negative number */
static void wnum(int width, int x) {
    char d[12];
    int n = 0, k;
    do { d[n++] = (char)('0' + x % 10); x /= 10; } while (x > 0);
    if (width > n) wpad((size_t)(width - n));
    for (k = n - 1; k >= 0; --k) wbuf[wlen++] = d[k];
}

void yb_maf_write(FILE *f, struct mafAli *a) {
    struct mafComp *c;
    int srcChars = 0, startChars = 0, sizeChars = 0, srcSizeChars = 0, row, n;
    size_t textLen, srcLen, nameLen;
    const char *dot, *chr;

    if (use_ref()) { ref_mafWrite(f, a); return; }
    wlen = 0;
    wneed(64);
    wbuf[wlen++] = 'a';
    if (a->score != MIN_INT) {
        n = snprintf(wbuf + wlen, wcap - wlen, " score=%3.1f", a->score);
        if ((size_t)n >= wcap - wlen) { wneed((size_t)n + 1); n = snprintf(wbuf + wlen, wcap - wlen, " score=%3.1f", a->score); }
        wlen += (size_t)n;
    }
    for (row = 0, c = a->components; c != NULL; c = c->next, row++) {
        switch (c->paralog) {
        case 's':
            break;
        case 'a':
            wneed(32);
            wlen += (size_t)sprintf(wbuf + wlen, " amplifier=%d", row);
            break;
        case 'c':
            wneed(32);
            wlen += (size_t)sprintf(wbuf + wlen, " copy=%d", row);
            break;
        default:
            fwrite(wbuf, 1, wlen, f);                      /* (the reference has printed this much by now) */
            fatalf("Wrong character: \'%c\'", c->paralog);
        }
    }
    wneed(2);
    wbuf[wlen++] = '\n';
    /* a negative number stops the reference in its width pass (digitsBaseTen), after the `a` line went out */
    for (c = a->components; c != NULL; c = c->next)
        if (c->start < 0 || c->size < 0 || c->srcSize < 0) { fwrite(wbuf, 1, wlen, f); wlen = 0; break; }
    for (c = a->components; c != NULL; c = c->next) {
        n = (int)strlen(c->src);
        if (n > srcChars) srcChars = n;
        n = digitsBaseTen(c->start); if (n > startChars) startChars = n;
        n = digitsBaseTen(c->size); if (n > sizeChars) sizeChars = n;
        n = digitsBaseTen(c->srcSize); if (n > srcSizeChars) srcSizeChars = n;
    }
    for (c = a->components; c != NULL; c = c->next) {
        /* parseSrcName + re-join (maf.c:283-288): name = up to the first '.', chr = what follows it (the name again if
         * nothing does); "name.chr" unless the two are equal */
        srcLen = strlen(c->src);
        dot = memchr(c->src, '.', srcLen);
        nameLen = dot ? (size_t)(dot - c->src) : srcLen;
        chr = (dot && dot[1]) ? dot + 1 : NULL;
        textLen = strlen(c->text);
        wneed(srcLen + textLen + (size_t)srcChars + 64);
        wbuf[wlen++] = 's'; wbuf[wlen++] = ' ';
        n = (int)wlen;
        memcpy(wbuf + wlen, c->src, nameLen); wlen += nameLen;
        if (chr != NULL && !(strlen(chr) == nameLen && memcmp(chr, c->src, nameLen) == 0)) {
            wbuf[wlen++] = '.';
            memcpy(wbuf + wlen, chr, srcLen - nameLen - 1); wlen += srcLen - nameLen - 1;
        }
        n = (int)wlen - n;
        if (n < srcChars) wpad((size_t)(srcChars - n));
        wbuf[wlen++] = ' ';
        wnum(startChars, c->start);
        wbuf[wlen++] = ' ';
        wnum(sizeChars, c->size);
        wbuf[wlen++] = ' ';
        wbuf[wlen++] = c->strand;
        wbuf[wlen++] = ' ';
        wnum(srcSizeChars, c->srcSize);
        wbuf[wlen++] = ' ';
        memcpy(wbuf + wlen, c->text, textLen); wlen += textLen;
        wbuf[wlen++] = '\n';
    }
    wneed(1);
    wbuf[wlen++] = '\n';
    fwrite(wbuf, 1, wlen, f);
}
