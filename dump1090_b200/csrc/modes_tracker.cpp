// modes_tracker.cpp — SURVEY.md §8(f) item 3: the per-aircraft state the reference keeps for its
// interactive table, HTTP map and SBS port (dump1090.c:1822-2164), as a small host library over
// the delivered message stream.  Pure host code: nothing here touches the GPU.
//
//   modes_tracker_update    interactiveReceiveData        dump1090.c:2069-2164
//   decode_airborne         decodeCPR                     :1952-1988 (global decode of an even/odd pair)
//   decode_surface          decodeCPRSurface              :2004-2052 (local decode against the reference position)
//   movement_knots          decodeMovementField           :2056-2066
//   modes_cpr_nl            cprNLFunction                 :1869-1931
//   modes_tracker_expire    interactiveRemoveStaleAircrafts :2205-2229
//   modes_tracker_format_json  aircraftsToJson            :2505-2551
//   modes_tracker_format_table interactiveShowData        :2167-2199
//   modes_format_sbs        modesSendSBSOutput            :2396-2446
//
// The arithmetic is kept expression for expression (doubles, the int truncations, the order of
// the operations), because positions are compared bit for bit with the reference's in the tests.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>
#include "modes_b200.h"

namespace {

// Transition latitudes of the "number of longitude zones" function NL (1090-WP-9-14): NL(lat) = k
// for lat below threshold[59-k] ...  Generated from the defining formula with NZ = 15 and rounded
// to the 8 decimals the published table carries; NL = 2 ends at exactly 87 degrees.
struct NlTable {
    double below[58];                                    // below[i]: |lat| < below[i]  ->  NL = 59 - i
    NlTable() {
        const double nz = 15.0;
        const double a = 1.0 - std::cos(M_PI / (2.0 * nz));
        for (int nl = 59; nl >= 2; nl--) {
            const double t = std::acos(std::sqrt(a / (1.0 - std::cos(2.0 * M_PI / nl)))) * 180.0 / M_PI;
            below[59 - nl] = std::round(t * 1e8) / 1e8;
        }
    }
};
const NlTable &nl_table() { static const NlTable t; return t; }

int cpr_nl(double lat) {
    if (lat < 0) lat = -lat;                             // symmetric about the equator
    const NlTable &t = nl_table();
    for (int i = 0; i < 58; i++)
        if (lat < t.below[i]) return 59 - i;
    return 1;
}

int cpr_mod(int a, int b) {                              // always non-negative
    int r = a % b;
    return r < 0 ? r + b : r;
}

int cpr_n(double lat, int is_odd) {
    int n = cpr_nl(lat) - is_odd;
    return n < 1 ? 1 : n;
}

double cpr_dlon(double lat, int is_odd) { return 360.0 / cpr_n(lat, is_odd); }

// Global decode of the aircraft's stored even/odd pair (17-bit fields, 2^17 = 131072).
// Leaves lat/lon untouched when the two frames straddle a latitude zone boundary.
void decode_airborne(modes_aircraft *a) {
    const double dlat_even = 360.0 / 60, dlat_odd = 360.0 / 59;
    const double lat_e = a->even_cprlat, lat_o = a->odd_cprlat;
    const double lon_e = a->even_cprlon, lon_o = a->odd_cprlon;

    const int j = (int)std::floor(((59 * lat_e - 60 * lat_o) / 131072) + 0.5);
    double rlat_e = dlat_even * (cpr_mod(j, 60) + lat_e / 131072);
    double rlat_o = dlat_odd * (cpr_mod(j, 59) + lat_o / 131072);
    if (rlat_e >= 270) rlat_e -= 360;
    if (rlat_o >= 270) rlat_o -= 360;
    if (cpr_nl(rlat_e) != cpr_nl(rlat_o)) return;

    if (a->even_cprtime > a->odd_cprtime) {              // the even frame is the newer one
        const int ni = cpr_n(rlat_e, 0);
        const int m = (int)std::floor((((lon_e * (cpr_nl(rlat_e) - 1)) - (lon_o * cpr_nl(rlat_e))) / 131072) + 0.5);
        a->lon = cpr_dlon(rlat_e, 0) * (cpr_mod(m, ni) + lon_e / 131072);
        a->lat = rlat_e;
    } else {
        const int ni = cpr_n(rlat_o, 1);
        const int m = (int)std::floor((((lon_e * (cpr_nl(rlat_o) - 1)) - (lon_o * cpr_nl(rlat_o))) / 131072.0) + 0.5);
        a->lon = cpr_dlon(rlat_o, 1) * (cpr_mod(m, ni) + lon_o / 131072);
        a->lat = rlat_o;
    }
    if (a->lon > 180) a->lon -= 360;
}

// Local decode of one surface position frame against the receiver's reference position (zones
// span 90 degrees on the ground, so a single frame is ambiguous without one).
void decode_surface(modes_aircraft *a, int fflag, int raw_lat, int raw_lon, double ref_lat, double ref_lon) {
    const double dlat = fflag ? 90.0 / 59 : 90.0 / 60;
    const int j = (int)std::floor(ref_lat / dlat) +
                  (int)std::floor(0.5 + cpr_mod((int)ref_lat, (int)dlat) / dlat - (double)raw_lat / 131072);
    double lat = dlat * (j + (double)raw_lat / 131072);
    if (std::fabs(lat - ref_lat) > 45) {                 // the other 90-degree solution is the near one
        if (lat > ref_lat) lat -= 90;
        else lat += 90;
    }
    if (lat < -90 || lat > 90) return;

    int ni = cpr_n(lat, fflag);
    if (ni == 0) ni = 1;
    const double dlon = 90.0 / ni;
    const int m = (int)std::floor(ref_lon / (90.0 / ni)) +
                  (int)std::floor(0.5 + cpr_mod((int)ref_lon, (int)(90.0 / ni)) / (90.0 / ni) - (double)raw_lon / 131072);
    double lon = dlon * (m + (double)raw_lon / 131072);
    while (lon > ref_lon + 45) lon -= 90;
    while (lon < ref_lon - 45) lon += 90;
    if (lon > 180) lon -= 360;
    if (lon < -180) lon += 360;
    a->lat = lat;
    a->lon = lon;
}

// Ground speed in knots from the 7-bit movement field; -1 = not available.  The quantisation
// steps widen with speed; fractions of a knot are dropped, as the reference's int return does.
int movement_knots(int movement) {
    if (movement == 0) return -1;
    if (movement == 1) return 0;
    if (movement <= 8) return (int)((movement - 2) * 0.125 + 0.125);
    if (movement <= 12) return (int)((movement - 9) * 0.25 + 1);
    if (movement <= 38) return (int)((movement - 13) * 0.5 + 2);
    if (movement <= 93) return (movement - 39) + 15;
    if (movement <= 108) return (movement - 94) * 2 + 70;
    if (movement <= 123) return (movement - 109) * 5 + 100;
    return 175;
}

// snprintf-style appender that keeps counting past the end of the buffer.
struct Out {
    char *buf; size_t cap, len = 0;
    Out(char *b, size_t c) : buf(b), cap(c) {}
    void add(const char *fmt, ...) {
        char tmp[320];
        va_list ap;
        va_start(ap, fmt);
        int n = vsnprintf(tmp, sizeof(tmp), fmt, ap);
        va_end(ap);
        if (n < 0) return;
        if ((size_t)n >= sizeof(tmp)) n = (int)sizeof(tmp) - 1;
        for (int i = 0; i < n; i++, len++)
            if (buf && len + 1 < cap) buf[len] = tmp[i];
    }
    size_t finish() {
        if (buf && cap) buf[len < cap ? len : cap - 1] = 0;
        return len;
    }
};

}  // namespace

struct modes_tracker {
    int check_crc = 1;
    std::deque<std::unique_ptr<modes_aircraft>> list;    // front = most recently created (the reference's list order)
    std::unordered_map<uint32_t, modes_aircraft *> by_addr;   // the reference walks its list; an index keeps busy skies O(1)
    double ref_lat = 0, ref_lon = 0;
    int ref_count = 0;
    modes_aircraft *find(uint32_t addr) {
        auto it = by_addr.find(addr);
        return it == by_addr.end() ? nullptr : it->second;
    }
};

extern "C" {

int modes_cpr_nl(double lat) { return cpr_nl(lat); }

modes_tracker *modes_tracker_create(int check_crc) {
    modes_tracker *t = new (std::nothrow) modes_tracker();
    if (t) t->check_crc = check_crc;
    return t;
}

void modes_tracker_destroy(modes_tracker *t) { delete t; }

const modes_aircraft *modes_tracker_update(modes_tracker *t, const modes_message *mm, int64_t now_ms) {
    if (!t || !mm) return nullptr;
    if (t->check_crc && mm->crcok == 0) return nullptr;
    const uint32_t addr = ((uint32_t)mm->aa1 << 16) | ((uint32_t)mm->aa2 << 8) | (uint32_t)mm->aa3;
    modes_aircraft *a = t->find(addr);
    if (!a) {
        std::unique_ptr<modes_aircraft> fresh(new (std::nothrow) modes_aircraft());
        if (!fresh) return nullptr;
        std::memset(fresh.get(), 0, sizeof(modes_aircraft));
        fresh->addr = addr;
        std::snprintf(fresh->hexaddr, sizeof(fresh->hexaddr), "%06x", (int)addr);
        a = fresh.get();
        t->by_addr[addr] = a;
        t->list.push_front(std::move(fresh));
    }
    a->seen = now_ms / 1000;
    a->messages++;

    const int df = mm->msgtype;
    if (df == 0 || df == 4 || df == 20) {
        a->altitude = mm->altitude;
    } else if (df == 17 || df == 18) {
        const int tc = mm->metype;
        if (tc >= 1 && tc <= 4) {
            std::memcpy(a->flight, mm->flight, sizeof(a->flight));
        } else if (tc >= 9 && tc <= 18) {
            a->altitude = mm->altitude;
            if (mm->fflag) {
                a->odd_cprlat = mm->raw_latitude; a->odd_cprlon = mm->raw_longitude; a->odd_cprtime = now_ms;
            } else {
                a->even_cprlat = mm->raw_latitude; a->even_cprlon = mm->raw_longitude; a->even_cprtime = now_ms;
            }
            if (std::llabs((long long)(a->even_cprtime - a->odd_cprtime)) <= 10000) {   // a pair at most 10 s apart
                const double prev_lat = a->lat, prev_lon = a->lon;
                decode_airborne(a);
                if (a->lat != prev_lat || a->lon != prev_lon) {
                    // a fresh airborne position also moves the receiver's reference position
                    if (t->ref_count == 0) {
                        t->ref_lat = a->lat; t->ref_lon = a->lon;
                    } else {
                        t->ref_lat += (a->lat - t->ref_lat) / (t->ref_count + 1);
                        t->ref_lon += (a->lon - t->ref_lon) / (t->ref_count + 1);
                    }
                    if (t->ref_count < 10000) t->ref_count++;
                }
            }
        } else if (tc >= 5 && tc <= 8) {
            if (t->ref_count) {
                if (mm->ground_track_valid) a->track = mm->ground_track;
                if (mm->movement_valid) a->speed = movement_knots(mm->movement);
                a->altitude = 0;
                decode_surface(a, mm->fflag, mm->raw_latitude, mm->raw_longitude, t->ref_lat, t->ref_lon);
            }
        } else if (tc == 19) {
            if (mm->mesub == 1 || mm->mesub == 2) {
                a->speed = mm->velocity;
                a->track = mm->heading;
            }
        }
    }
    return a;
}

size_t modes_tracker_count(const modes_tracker *t) { return t ? t->list.size() : 0; }

size_t modes_tracker_list(const modes_tracker *t, modes_aircraft *out, size_t capacity) {
    if (!t) return 0;
    size_t n = 0;
    for (const auto &a : t->list) {
        if (out && n < capacity) out[n] = *a;
        n++;
    }
    return n;
}

size_t modes_tracker_expire(modes_tracker *t, int64_t now_ms, int ttl_seconds) {
    if (!t) return 0;
    const int64_t now = now_ms / 1000;
    size_t removed = 0;
    for (auto it = t->list.begin(); it != t->list.end();) {
        if ((now - (*it)->seen) > ttl_seconds) { t->by_addr.erase((*it)->addr); it = t->list.erase(it); removed++; }
        else ++it;
    }
    return removed;
}

void modes_tracker_reference(const modes_tracker *t, double *lat, double *lon, int *count) {
    if (lat) *lat = t ? t->ref_lat : 0;
    if (lon) *lon = t ? t->ref_lon : 0;
    if (count) *count = t ? t->ref_count : 0;
}

size_t modes_tracker_format_json(const modes_tracker *t, int metric, char *buf, size_t capacity) {
    Out o(buf, capacity);
    o.add("[\n");
    bool any = false;
    if (t) {
        for (const auto &a : t->list) {
            int altitude = a->altitude, speed = a->speed;
            if (metric) { altitude = (int)(altitude / 3.2828); speed = (int)(speed * 1.852); }
            if (a->lat != 0 && a->lon != 0) {
                if (any) o.add(",\n");
                o.add("{\"hex\":\"%s\", \"flight\":\"%s\", \"lat\":%f, \"lon\":%f, \"altitude\":%d, \"track\":%d, \"speed\":%d}",
                      a->hexaddr, a->flight, a->lat, a->lon, altitude, a->track, speed);
                any = true;
            }
        }
    }
    if (any) o.add("\n");
    o.add("]\n");
    return o.finish();
}

size_t modes_tracker_format_table(const modes_tracker *t, int metric, int max_rows, int64_t now_ms, char *buf, size_t capacity) {
    Out o(buf, capacity);
    const int64_t now = now_ms / 1000;
    char progress[4] = {' ', ' ', ' ', 0};
    progress[now % 3] = '.';
    o.add("\x1b[H\x1b[2J");
    o.add("Hex    Flight   Altitude  Speed   Lat       Lon       Track  Messages Seen %s\n"
          "--------------------------------------------------------------------------------\n", progress);
    int rows = 0;
    if (t) {
        for (const auto &a : t->list) {
            if (rows >= max_rows) break;
            int altitude = a->altitude, speed = a->speed;
            if (metric) { altitude = (int)(altitude / 3.2828); speed = (int)(speed * 1.852); }
            o.add("%-6s %-8s %-9d %-7d %-7.03f   %-7.03f   %-3d   %-9ld %d sec\n", a->hexaddr, a->flight, altitude, speed,
                  a->lat, a->lon, a->track, (long)a->messages, (int)(now - a->seen));
            rows++;
        }
    }
    return o.finish();
}

size_t modes_format_sbs(const modes_message *mm, const modes_aircraft *a, char *buf, size_t capacity) {
    if (!mm || !a) return 0;
    Out o(buf, capacity);
    int emergency = 0, ground = 0, alert = 0, spi = 0;
    const int df = mm->msgtype;
    if (df == 4 || df == 5 || df == 21) {
        // the squawk is kept as four octal digits read in base 10
        if (mm->identity == 7500 || mm->identity == 7600 || mm->identity == 7700) emergency = -1;
        if (mm->fs == 1 || mm->fs == 3) ground = -1;
        if (mm->fs == 2 || mm->fs == 3 || mm->fs == 4) alert = -1;
        if (mm->fs == 4 || mm->fs == 5) spi = -1;
    }
    const int a1 = mm->aa1, a2 = mm->aa2, a3 = mm->aa3;
    const bool es = df == 17 || df == 18;
    if (df == 0) {
        o.add("MSG,5,,,%02X%02X%02X,,,,,,,%d,,,,,,,,,,", a1, a2, a3, mm->altitude);
    } else if (df == 4) {
        o.add("MSG,5,,,%02X%02X%02X,,,,,,,%d,,,,,,,%d,%d,%d,%d", a1, a2, a3, mm->altitude, alert, emergency, spi, ground);
    } else if (df == 5 || df == 21) {
        o.add("MSG,6,,,%02X%02X%02X,,,,,,,,,,,,,%d,%d,%d,%d,%d", a1, a2, a3, mm->identity, alert, emergency, spi, ground);
    } else if (df == 11) {
        o.add("MSG,8,,,%02X%02X%02X,,,,,,,,,,,,,,,,,", a1, a2, a3);
    } else if (es && mm->metype == 4) {
        o.add("MSG,1,,,%02X%02X%02X,,,,,,%s,,,,,,,,0,0,0,0", a1, a2, a3, mm->flight);
    } else if (es && mm->metype >= 9 && mm->metype <= 18) {
        if (a->lat == 0 && a->lon == 0)
            o.add("MSG,3,,,%02X%02X%02X,,,,,,,%d,,,,,,,0,0,0,0", a1, a2, a3, mm->altitude);
        else
            o.add("MSG,3,,,%02X%02X%02X,,,,,,,%d,,,%1.5f,%1.5f,,,0,0,0,0", a1, a2, a3, mm->altitude, a->lat, a->lon);
    } else if (es && mm->metype == 19 && mm->mesub == 1) {
        const int vr = (mm->vert_rate_sign == 0 ? 1 : -1) * (mm->vert_rate - 1) * 64;
        o.add("MSG,4,,,%02X%02X%02X,,,,,,,,%d,%d,,,%i,,0,0,0,0", a1, a2, a3, a->speed, a->track, vr);
    } else {
        if (buf && capacity) buf[0] = 0;
        return 0;
    }
    o.add("\n");
    return o.finish();
}

}  // extern "C"
