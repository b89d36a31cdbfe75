/* Stand-in for librtlsdr's <rtl-sdr.h>, which is not installed in this image.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference translation
 * unit includes "rtl-sdr.h" (dump1090.c:46) and calls 15 librtlsdr functions
 * from its live-radio path (dump1090.c:390-433, :520, :3008).  None of them is
 * reached by the --ifile decode path the oracle exercises, so every function
 * here is a no-op that reports "no device".  Signatures follow the public
 * librtlsdr API so the reference compiles unmodified.
 */
#ifndef ORACLE_STUB_RTL_SDR_H
#define ORACLE_STUB_RTL_SDR_H
#include <stdint.h>

typedef struct rtlsdr_dev rtlsdr_dev_t;
typedef void (*rtlsdr_read_async_cb_t)(unsigned char *buf, uint32_t len, void *ctx);

static inline uint32_t rtlsdr_get_device_count(void) { return 0; }
static inline int rtlsdr_get_device_usb_strings(uint32_t i, char *m, char *p, char *s)
{ (void)i; if (m) m[0] = 0; if (p) p[0] = 0; if (s) s[0] = 0; return -1; }
static inline int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index) { (void)index; *dev = 0; return -1; }
static inline int rtlsdr_close(rtlsdr_dev_t *dev) { (void)dev; return 0; }
static inline int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int manual) { (void)dev; (void)manual; return -1; }
static inline int rtlsdr_get_tuner_gains(rtlsdr_dev_t *dev, int *gains) { (void)dev; (void)gains; return 0; }
static inline int rtlsdr_set_tuner_gain(rtlsdr_dev_t *dev, int gain) { (void)dev; (void)gain; return -1; }
static inline int rtlsdr_get_tuner_gain(rtlsdr_dev_t *dev) { (void)dev; return 0; }
static inline int rtlsdr_set_freq_correction(rtlsdr_dev_t *dev, int ppm) { (void)dev; (void)ppm; return -1; }
static inline int rtlsdr_set_agc_mode(rtlsdr_dev_t *dev, int on) { (void)dev; (void)on; return -1; }
static inline int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq) { (void)dev; (void)freq; return -1; }
static inline int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate) { (void)dev; (void)rate; return -1; }
static inline int rtlsdr_reset_buffer(rtlsdr_dev_t *dev) { (void)dev; return -1; }
static inline int rtlsdr_read_async(rtlsdr_dev_t *dev, rtlsdr_read_async_cb_t cb, void *ctx,
                                    uint32_t buf_num, uint32_t buf_len)
{ (void)dev; (void)cb; (void)ctx; (void)buf_num; (void)buf_len; return -1; }
#endif
