#define GX_KW 2
#include "gx_kw.inl"
