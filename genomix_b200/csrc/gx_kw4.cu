#define GX_KW 4
#include "gx_kw.inl"
