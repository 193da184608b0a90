/* declarations of oracle/ref_cuda_shim.c for the reference's main (force-included by oracle/Makefile:ref_cuda). Test infrastructure. */
void emsar_shim_bind(void);
void emsar_shim_counts(void);
void emsar_shim_estimate(void);
