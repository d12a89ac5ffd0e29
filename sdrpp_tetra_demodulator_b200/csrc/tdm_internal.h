// tdm_internal.h -- shared between the translation units of libtdm_b200.so; not part of the ABI.
#pragma once
// Records the text tdm_last_error() returns (thread local) and hands `code` back.
int tdm_internal_fail(int code, const char* fmt, ...);
