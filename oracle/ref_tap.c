/* oracle/ref_tap.c — strong replacement for the reference's message sink.
 * TEST INFRASTRUCTURE ONLY.  Linked after `objcopy --weaken-symbol
 * useModesMessage` on the reference object so detectModeS()'s call at
 * dump1090.c:1777 lands here; we forward to the harness, which knows the
 * struct layout. */
void ref_harness_sink(void *mm);
void useModesMessage(void *mm) { ref_harness_sink(mm); }
