/* oracle/ref_harness.c — single-threaded driver around the UNMODIFIED reference.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product; only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
 *
 * This translation unit textually includes /root/reference/dump1090.c (never
 * copied into this repo; the build recipe passes -I/root/reference and
 * -Dmain=ref_main) and drives the reference's own functions one buffer at a
 * time, exactly as its main loop does (dump1090.c:2968-2990) but without the
 * reader thread, so the EOF-buffer race (dump1090.c:497 vs :2989) becomes an
 * explicit switch:
 *
 *   drop_eof = 0   every buffer is decoded            (284 lines on modes1.bin)
 *   drop_eof = 1   the buffer that hit EOF is dropped (217 lines; what the
 *                  stock binary prints in most runs)
 *
 * Buffer feeding restates readDataFromFile (dump1090.c:481-507): 476-byte
 * carry memcpy, up to 262144 new bytes, 127-padding of a short read.
 *
 * The message sink: the reference calls useModesMessage() (dump1090.c:1777).
 * The build recipe weakens that symbol in this object (objcopy
 * --weaken-symbol) and links ref_tap.c, whose strong useModesMessage()
 * forwards to ref_harness_sink() below.  The sink re-applies the reference's
 * gate (dump1090.c:1803) and then records the struct and/or lets the
 * reference's own displayModesMessage() print it.
 *
 * time() is frozen (the recipe passes -Dtime=ref_frozen_time) so the 60 s
 * ICAO-cache TTL (dump1090.c:913, :924) cannot fire in long timing runs.
 */
#include <time.h>
time_t ref_frozen_time(time_t *t);

#include "dump1090.c"           /* resolved through -I/root/reference */

#include "oracle_msg.h"

/* The clock the reference sees: frozen at 10^9 s for the decode tests, set per message by the
 * tracker door below (time() and gettimeofday() are both redirected by the build recipe). */
#define REF_CLOCK_DEFAULT_MS 1000000000000LL
static long long g_clock_ms = REF_CLOCK_DEFAULT_MS;
time_t ref_frozen_time(time_t *t) {
    time_t v = (time_t)(g_clock_ms / 1000);
    if (t) *t = v;
    return v;
}
int ref_gettimeofday(struct timeval *tv, void *tz) {
    (void)tz;
    tv->tv_sec = g_clock_ms / 1000;
    tv->tv_usec = (g_clock_ms % 1000) * 1000;
    return 0;
}

static struct oracle_msg *g_out = NULL;
static size_t g_out_cap = 0, g_out_n = 0;
static int g_print = 0;
static int g_inited = 0;

static void copy_fields(struct oracle_msg *o, const struct modesMessage *mm) {
    memset(o, 0, sizeof(*o));
    memcpy(o->msg, mm->msg, 14);
    o->msgbits = mm->msgbits; o->msgtype = mm->msgtype; o->crcok = mm->crcok;
    o->crc = mm->crc; o->errorbit = mm->errorbit;
    o->aa1 = mm->aa1; o->aa2 = mm->aa2; o->aa3 = mm->aa3;
    o->phase_corrected = mm->phase_corrected;
    o->ca = mm->ca; o->iid = mm->iid; o->metype = mm->metype; o->mesub = mm->mesub;
    o->heading_is_valid = mm->heading_is_valid; o->heading = mm->heading;
    o->aircraft_type = mm->aircraft_type; o->fflag = mm->fflag; o->tflag = mm->tflag;
    o->raw_latitude = mm->raw_latitude; o->raw_longitude = mm->raw_longitude;
    memcpy(o->flight, mm->flight, 9);
    o->ew_dir = mm->ew_dir; o->ew_velocity = mm->ew_velocity;
    o->ns_dir = mm->ns_dir; o->ns_velocity = mm->ns_velocity;
    o->vert_rate_source = mm->vert_rate_source; o->vert_rate_sign = mm->vert_rate_sign;
    o->vert_rate = mm->vert_rate; o->velocity = mm->velocity;
    o->movement = mm->movement; o->movement_valid = mm->movement_valid;
    o->ground_track = mm->ground_track; o->ground_track_valid = mm->ground_track_valid;
    o->fs = mm->fs; o->dr = mm->dr; o->um = mm->um; o->identity = mm->identity;
    o->altitude = mm->altitude; o->unit = mm->unit;
    o->sample_pos = -1;         /* not observable from outside detectModeS */
}

/* Called by ref_tap.c's useModesMessage().  Gate restated from dump1090.c:1803. */
void ref_harness_sink(void *p) {
    struct modesMessage *mm = (struct modesMessage *)p;
    if (Modes.stats) return;
    if (!(Modes.check_crc == 0 || mm->crcok)) return;
    if (g_out && g_out_n < g_out_cap) copy_fields(&g_out[g_out_n], mm);
    g_out_n++;
    if (g_print) {
        displayModesMessage(mm);                               /* :1812 */
        if (!Modes.raw && !Modes.onlyaddr) printf("\n");       /* :1813 */
    }
}

static void ref_reset(int fix_errors, int aggressive, int check_crc, int stats) {
    if (!g_inited) {
        modesInitConfig();
        modesInit();
        g_inited = 1;
    }
    Modes.fix_errors = fix_errors;
    Modes.aggressive = aggressive;
    Modes.check_crc = check_crc;
    Modes.stats = stats;
    Modes.interactive = 0;
    Modes.net = 0;
    g_clock_ms = REF_CLOCK_DEFAULT_MS;
    memset(Modes.data, 127, Modes.data_len);                                   /* :344 */
    memset(Modes.icao_cache, 0, sizeof(uint32_t) * MODES_ICAO_CACHE_LEN * 2);  /* :336 */
    Modes.stat_valid_preamble = 0; Modes.stat_demodulated = 0;
    Modes.stat_goodcrc = 0; Modes.stat_badcrc = 0; Modes.stat_fixed = 0;
    Modes.stat_single_bit_fix = 0; Modes.stat_two_bits_fix = 0;
    Modes.stat_out_of_phase = 0;
}

/* Feed the byte stream through the reference decode path.  Returns the number
 * of buffers decoded. */
static long ref_feed(const unsigned char *iq, size_t nbytes, int drop_eof) {
    size_t off = 0;
    long nbuf = 0;
    for (;;) {
        size_t avail = nbytes - off;
        size_t take = avail < MODES_DATA_LEN ? avail : MODES_DATA_LEN;
        int eof = take < MODES_DATA_LEN;
        memcpy(Modes.data, Modes.data + MODES_DATA_LEN, (MODES_FULL_LEN - 1) * 4);  /* :481 */
        memcpy(Modes.data + (MODES_FULL_LEN - 1) * 4, iq + off, take);
        if (eof)
            memset(Modes.data + (MODES_FULL_LEN - 1) * 4 + take, 127, MODES_DATA_LEN - take); /* :506 */
        off += take;
        if (eof && drop_eof) break;
        computeMagnitudeVector();                                /* :2974 */
        detectModeS(Modes.magnitude, Modes.data_len / 2);        /* :2986 */
        nbuf++;
        if (eof) break;
    }
    return nbuf;
}

static void ref_get_stats(long long *st) {
    st[0] = Modes.stat_valid_preamble; st[1] = Modes.stat_out_of_phase;
    st[2] = Modes.stat_demodulated;    st[3] = Modes.stat_goodcrc;
    st[4] = Modes.stat_badcrc;         st[5] = Modes.stat_fixed;
    st[6] = Modes.stat_single_bit_fix; st[7] = Modes.stat_two_bits_fix;
}

/* ctypes entry: decode a whole stream, collect messages.  Returns message
 * count (may exceed cap; only cap are stored).  stats[8] as in ref_get_stats. */
long ref_decode(const unsigned char *iq, size_t nbytes, int fix_errors, int aggressive,
                int check_crc, int drop_eof, struct oracle_msg *out, size_t cap,
                long long *stats) {
    ref_reset(fix_errors, aggressive, check_crc, 0);
    g_out = out; g_out_cap = cap; g_out_n = 0; g_print = 0;
    ref_feed(iq, nbytes, drop_eof);
    if (stats) ref_get_stats(stats);
    g_out = NULL;
    return (long)g_out_n;
}

/* ctypes entry: timing.  Runs the reference loop with Modes.stats=1 (sink
 * silent, dump1090.c:1803) `loops` times over the stream and returns seconds
 * of wall-clock for the decode loops only. */
double ref_time_decode(const unsigned char *iq, size_t nbytes, int fix_errors, int aggressive,
                       int check_crc, int loops, long long *stats) {
    struct timespec a, b;
    ref_reset(fix_errors, aggressive, check_crc, 1);
    g_out = NULL; g_print = 0;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int i = 0; i < loops; i++) ref_feed(iq, nbytes, 0);
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (stats) ref_get_stats(stats);
    return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}

/* ctypes entry: the same loop with the two hot calls timed separately (report
 * computeMagnitudeVector and detectModeS separately and combined).  seconds[0] = magnitude,
 * seconds[1] = detectModeS, summed over all buffers of `loops` passes. */
void ref_time_phases(const unsigned char *iq, size_t nbytes, int fix_errors, int aggressive,
                     int check_crc, int loops, double seconds[2]) {
    struct timespec a, b, c;
    ref_reset(fix_errors, aggressive, check_crc, 1);
    g_out = NULL; g_print = 0;
    seconds[0] = seconds[1] = 0;
    for (int i = 0; i < loops; i++) {
        size_t off = 0;
        for (;;) {                                                   /* ref_feed, instrumented */
            size_t avail = nbytes - off;
            size_t take = avail < MODES_DATA_LEN ? avail : MODES_DATA_LEN;
            int eof = take < MODES_DATA_LEN;
            memcpy(Modes.data, Modes.data + MODES_DATA_LEN, (MODES_FULL_LEN - 1) * 4);
            memcpy(Modes.data + (MODES_FULL_LEN - 1) * 4, iq + off, take);
            if (eof) memset(Modes.data + (MODES_FULL_LEN - 1) * 4 + take, 127, MODES_DATA_LEN - take);
            off += take;
            clock_gettime(CLOCK_MONOTONIC, &a);
            computeMagnitudeVector();
            clock_gettime(CLOCK_MONOTONIC, &b);
            detectModeS(Modes.magnitude, Modes.data_len / 2);
            clock_gettime(CLOCK_MONOTONIC, &c);
            seconds[0] += (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
            seconds[1] += (c.tv_sec - b.tv_sec) + 1e-9 * (c.tv_nsec - b.tv_nsec);
            if (eof) break;
        }
    }
}

/* ---- tracker door: the reference's interactiveReceiveData / decodeCPR /
 * modesSendSBSOutput / aircraftsToJson driven message by message with an explicit clock. */
static char g_sbs[512];
static int g_sbs_len = 0;
void ref_harness_net_sink(int service, void *msg, int len) {
    (void)service;
    if (len > (int)sizeof(g_sbs) - 1) len = (int)sizeof(g_sbs) - 1;
    memcpy(g_sbs, msg, len);
    g_sbs_len = len;
}

static void copy_aircraft(struct oracle_aircraft *o, const struct aircraft *a) {
    memset(o, 0, sizeof(*o));
    o->addr = a->addr;
    memcpy(o->hexaddr, a->hexaddr, 7);
    memcpy(o->flight, a->flight, 9);
    o->altitude = a->altitude; o->speed = a->speed; o->track = a->track;
    o->seen = a->seen; o->messages = a->messages;
    o->odd_cprlat = a->odd_cprlat; o->odd_cprlon = a->odd_cprlon;
    o->even_cprlat = a->even_cprlat; o->even_cprlon = a->even_cprlon;
    o->lat = a->lat; o->lon = a->lon;
    o->odd_cprtime = a->odd_cprtime; o->even_cprtime = a->even_cprtime;
}

void ref_track_reset(int check_crc) {
    ref_reset(1, 0, check_crc, 0);
    while (Modes.aircrafts) { struct aircraft *n = Modes.aircrafts->next; free(Modes.aircrafts); Modes.aircrafts = n; }
    Modes.ref_lat = Modes.ref_lon = 0;
    Modes.ref_count = 0;
    Modes.sbsos = 7;                    /* any value: the fan-out is intercepted */
}

/* Returns 1 and fills *out (and sbs: the SBS line, "" when the message type has none) when the
 * reference tracked the message, 0 when it ignored it. */
int ref_track_update(const struct oracle_msg *m, long long now_ms, struct oracle_aircraft *out, char *sbs512) {
    struct modesMessage mm;
    memset(&mm, 0, sizeof(mm));
    memcpy(mm.msg, m->msg, 14);
    mm.msgbits = m->msgbits; mm.msgtype = m->msgtype; mm.crcok = m->crcok; mm.crc = m->crc;
    mm.errorbit = m->errorbit; mm.aa1 = m->aa1; mm.aa2 = m->aa2; mm.aa3 = m->aa3;
    mm.phase_corrected = m->phase_corrected; mm.ca = m->ca; mm.iid = m->iid;
    mm.metype = m->metype; mm.mesub = m->mesub; mm.heading_is_valid = m->heading_is_valid;
    mm.heading = m->heading; mm.aircraft_type = m->aircraft_type; mm.fflag = m->fflag; mm.tflag = m->tflag;
    mm.raw_latitude = m->raw_latitude; mm.raw_longitude = m->raw_longitude;
    memcpy(mm.flight, m->flight, 9);
    mm.ew_dir = m->ew_dir; mm.ew_velocity = m->ew_velocity; mm.ns_dir = m->ns_dir; mm.ns_velocity = m->ns_velocity;
    mm.vert_rate_source = m->vert_rate_source; mm.vert_rate_sign = m->vert_rate_sign;
    mm.vert_rate = m->vert_rate; mm.velocity = m->velocity;
    mm.movement = m->movement; mm.movement_valid = m->movement_valid;
    mm.ground_track = m->ground_track; mm.ground_track_valid = m->ground_track_valid;
    mm.fs = m->fs; mm.dr = m->dr; mm.um = m->um; mm.identity = m->identity;
    mm.altitude = m->altitude; mm.unit = m->unit;
    g_clock_ms = now_ms;
    struct aircraft *a = interactiveReceiveData(&mm);
    if (sbs512) sbs512[0] = 0;
    if (!a) return 0;
    g_sbs_len = 0;
    modesSendSBSOutput(&mm, a);
    if (sbs512) { memcpy(sbs512, g_sbs, g_sbs_len); sbs512[g_sbs_len] = 0; }
    copy_aircraft(out, a);
    return 1;
}

long ref_track_list(struct oracle_aircraft *out, long cap) {
    long n = 0;
    for (struct aircraft *a = Modes.aircrafts; a; a = a->next, n++)
        if (n < cap) copy_aircraft(out + n, a);
    return n;
}

long ref_track_expire(long long now_ms, int ttl_seconds) {
    long before = 0, after = 0;
    for (struct aircraft *a = Modes.aircrafts; a; a = a->next) before++;
    g_clock_ms = now_ms;
    Modes.interactive_ttl = ttl_seconds;
    interactiveRemoveStaleAircrafts();
    for (struct aircraft *a = Modes.aircrafts; a; a = a->next) after++;
    return before - after;
}

int ref_track_json(int metric, char *buf, int cap) {
    int len = 0;
    Modes.metric = metric;
    char *c = aircraftsToJson(&len);
    Modes.metric = 0;
    int n = len < cap - 1 ? len : cap - 1;
    memcpy(buf, c, n);
    buf[n] = 0;
    free(c);
    return len;
}

/* interactiveShowData() prints to stdout: borrow fd 1 for the call and read the text back. */
int ref_track_table(int metric, int max_rows, long long now_ms, char *buf, int cap) {
    char path[] = "/tmp/ref_table_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) return -1;
    fflush(stdout);
    int saved = dup(1);
    dup2(fd, 1);
    g_clock_ms = now_ms;
    Modes.metric = metric;
    Modes.interactive_rows = max_rows;
    interactiveShowData();
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    Modes.metric = 0;
    off_t len = lseek(fd, 0, SEEK_END);
    lseek(fd, 0, SEEK_SET);
    int n = (int)(len < cap - 1 ? len : cap - 1);
    if (read(fd, buf, n) != n) n = 0;
    buf[n] = 0;
    close(fd);
    unlink(path);
    return (int)len;
}

void ref_track_reference(double *lat, double *lon, int *count) {
    *lat = Modes.ref_lat; *lon = Modes.ref_lon; *count = Modes.ref_count;
}

int ref_cpr_nl(double lat) { return cprNLFunction(lat); }

/* ctypes entry: the two reference kernels on one caller-supplied buffer image
 * (262620 bytes, carry included) — lets tests pin the magnitude vector. */
void ref_magnitude(const unsigned char *data262620, uint16_t *mag131310) {
    ref_reset(1, 0, 1, 1);
    memcpy(Modes.data, data262620, Modes.data_len);
    computeMagnitudeVector();
    memcpy(mag131310, Modes.magnitude, Modes.data_len);   /* data_len/2 u16 = data_len bytes */
}

uint32_t ref_checksum(const unsigned char *msg, int bits) {
    unsigned char tmp[14];
    memcpy(tmp, msg, 14);
    return modesChecksum(tmp, bits);
}

/* ctypes entry: decodeModesMessage on raw bytes with a fresh ICAO cache
 * (hex known-answer door, like decodeHexMessage dump1090.c:2472-2502). */
void ref_decode_bytes(const unsigned char *msg14, int fix_errors, int aggressive,
                      struct oracle_msg *out) {
    struct modesMessage mm;
    unsigned char tmp[14];
    ref_reset(fix_errors, aggressive, 1, 1);
    memcpy(tmp, msg14, 14);
    memset(&mm, 0, sizeof(mm));
    decodeModesMessage(&mm, tmp);
    copy_fields(out, &mm);
}

/* ctypes entry: the reference's own raw-TCP input handler (decodeHexMessage, dump1090.c:2472)
 * on one text line, fresh ICAO cache.  Returns 1 and fills *out if it delivered a message
 * (check_crc off, so every parsed line delivers), 0 if it discarded the line. */
int ref_decode_hex_line(const char *line, int fix_errors, int aggressive, struct oracle_msg *out) {
    static struct client c;
    ref_reset(fix_errors, aggressive, 0, 0);
    memset(&c, 0, sizeof(c));
    strncpy(c.buf, line, MODES_CLIENT_BUF_SIZE);
    g_out = out; g_out_cap = 1; g_out_n = 0; g_print = 0;
    decodeHexMessage(&c);
    g_out = NULL;
    return g_out_n > 0;
}

#ifdef REF_HARNESS_MAIN
/* CLI: ref_dump1090 --ifile F [--raw] [--no-fix] [--aggressive] [--no-crc-check]
 *                   [--stats] [--onlyaddr] [--drop-eof-buffer] [--time LOOPS] */
#undef main
int main(int argc, char **argv) {
    const char *fn = NULL;
    int fix = 1, aggr = 0, crc = 1, stats = 0, drop = 0, loops = 0;
    modesInitConfig(); modesInit(); g_inited = 1;
    for (int j = 1; j < argc; j++) {
        if (!strcmp(argv[j], "--ifile") && j + 1 < argc) fn = argv[++j];
        else if (!strcmp(argv[j], "--raw")) Modes.raw = 1;
        else if (!strcmp(argv[j], "--onlyaddr")) Modes.onlyaddr = 1;
        else if (!strcmp(argv[j], "--no-fix")) fix = 0;
        else if (!strcmp(argv[j], "--aggressive")) aggr = 1;
        else if (!strcmp(argv[j], "--no-crc-check")) crc = 0;
        else if (!strcmp(argv[j], "--stats")) stats = 1;
        else if (!strcmp(argv[j], "--drop-eof-buffer")) drop = 1;
        else if (!strcmp(argv[j], "--time") && j + 1 < argc) loops = atoi(argv[++j]);
        else { fprintf(stderr, "unknown option %s\n", argv[j]); return 1; }
    }
    if (!fn) { fprintf(stderr, "need --ifile\n"); return 1; }
    FILE *f = fopen(fn, "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    unsigned char *buf = malloc(sz ? sz : 1);
    if (fread(buf, 1, sz, f) != (size_t)sz) { perror("read"); return 1; }
    fclose(f);
    if (loops > 0) {
        long long st[8];
        double s = ref_time_decode(buf, sz, fix, aggr, crc, loops, st);
        printf("%.6f s for %d loops of %ld samples: %.2f Msamples/s\n", s, loops, sz / 2,
               1e-6 * loops * (sz / 2) / s);
        return 0;
    }
    ref_reset(fix, aggr, crc, stats);
    g_print = 1;
    ref_feed(buf, sz, drop);
    if (stats) {   /* same report as dump1090.c:2994-3005 */
        printf("%lld valid preambles\n", Modes.stat_valid_preamble);
        printf("%lld demodulated again after phase correction\n", Modes.stat_out_of_phase);
        printf("%lld demodulated with zero errors\n", Modes.stat_demodulated);
        printf("%lld with good crc\n", Modes.stat_goodcrc);
        printf("%lld with bad crc\n", Modes.stat_badcrc);
        printf("%lld errors corrected\n", Modes.stat_fixed);
        printf("%lld single bit errors\n", Modes.stat_single_bit_fix);
        printf("%lld two bits errors\n", Modes.stat_two_bits_fix);
        printf("%lld total usable messages\n", Modes.stat_goodcrc + Modes.stat_fixed);
    }
    return 0;
}
#endif
