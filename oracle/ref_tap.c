/* oracle/ref_tap.c — strong replacement for the reference's message sink.
 * TEST INFRASTRUCTURE ONLY.  Linked after `objcopy --weaken-symbol
 * useModesMessage` on the reference object so detectModeS()'s call at
 * dump1090.c:1777 lands here; we forward to the harness, which knows the
 * struct layout. */
void ref_harness_sink(void *mm);
void useModesMessage(void *mm) { ref_harness_sink(mm); }

/* Same trick for the network fan-out (dump1090.c:2346): the harness captures what
 * modesSendSBSOutput() would have written to its clients. */
void ref_harness_net_sink(int service, void *msg, int len);
void modesSendAllClients(int service, void *msg, int len) { ref_harness_net_sink(service, msg, len); }
