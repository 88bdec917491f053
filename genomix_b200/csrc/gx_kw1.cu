#define GX_KW 1
#include "gx_kw.inl"
