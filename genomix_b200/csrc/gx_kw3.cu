#define GX_KW 3
#include "gx_kw.inl"
