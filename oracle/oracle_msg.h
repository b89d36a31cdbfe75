/* oracle/oracle_msg.h — flat message record shared by the reference harness
 * (ref_harness.c) and the CPU restatement (modes_oracle.c).  TEST
 * INFRASTRUCTURE ONLY.  Field names and meaning follow struct modesMessage
 * (dump1090.c:211-260); the layout (all 32-bit, fixed order) is chosen so one
 * ctypes.Structure in tests/ reads records from the reference, the oracle and
 * the product (include/modes_b200.h declares the same layout independently).
 */
#ifndef ORACLE_MSG_H
#define ORACLE_MSG_H
#include <stdint.h>

struct oracle_msg {
    uint8_t  msg[14];           /* binary message (after any CRC fix) */
    uint8_t  pad0[2];
    int32_t  msgbits, msgtype, crcok;
    uint32_t crc;
    int32_t  errorbit, aa1, aa2, aa3, phase_corrected;
    int32_t  ca, iid;
    int32_t  metype, mesub, heading_is_valid, heading, aircraft_type;
    int32_t  fflag, tflag, raw_latitude, raw_longitude;
    char     flight[9];
    char     pad1[3];
    int32_t  ew_dir, ew_velocity, ns_dir, ns_velocity;
    int32_t  vert_rate_source, vert_rate_sign, vert_rate, velocity;
    int32_t  movement, movement_valid, ground_track, ground_track_valid;
    int32_t  fs, dr, um, identity;
    int32_t  altitude, unit;
    int32_t  nfixed;            /* extra: bits corrected (0/1/2); reference harness leaves 0 */
    int32_t  pad2;
    int64_t  sample_pos;        /* extra: stream sample index of the preamble start, -1 if unknown */
};

/* struct aircraft (dump1090.c:112-130) as the tracker tests read it back; same layout as
 * modes_aircraft in include/modes_b200.h. */
struct oracle_aircraft {
    uint32_t addr;
    char     hexaddr[7];
    char     flight[9];
    int32_t  altitude, speed, track;
    int64_t  seen;
    int64_t  messages;
    int32_t  odd_cprlat, odd_cprlon, even_cprlat, even_cprlon;
    double   lat, lon;
    int64_t  odd_cprtime, even_cprtime;
};
#endif
