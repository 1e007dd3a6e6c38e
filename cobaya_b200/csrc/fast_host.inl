// fast_host.inl -- host-side packing for the fast kernels (included by engine.cu)
static int pack_fast(cb2_engine *) { return 0; }
