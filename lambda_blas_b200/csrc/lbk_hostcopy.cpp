// lbk_hostcopy.cpp -- the host-side copy of the staged transfers (lbk_lib.cu, staged_transfer).
//
// A synchronous lbk_vec_upload / lbk_vec_download from PAGEABLE memory (what a Data.Vector.Storable buffer is, Single.hs:67) moves
// every byte twice: once between the caller's buffer and a pinned staging chunk, on the CPU, and once over PCIe.  Several lanes
// (threads) do the CPU half at once, and with them the host's memory bandwidth becomes the limit: a plain memcpy of an 8 MiB chunk
// reads the source, reads the DESTINATION line as well (write-allocate) and writes it back -- three transfers for one.  The
// destination is not going to be read by this core again (the copy engine, or the caller much later, reads it from memory), so it is
// written with non-temporal stores: two transfers for one.  glibc does the same, but only above a threshold (three quarters of the
// shared cache) that an 8 MiB chunk does not reach.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>

extern "C" void lbk_host_stream_copy(void *dst, const void *src, size_t bytes)
{
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    if (bytes < 4096) { memcpy(d, s, bytes); return; }
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;      // non-temporal stores want an aligned destination
    if (head) { memcpy(d, s, head); d += head; s += head; bytes -= head; }
    const size_t body = bytes & ~static_cast<size_t>(63);
    for (size_t i = 0; i < body; i += 64) {
        _mm_prefetch(s + i + 1024, _MM_HINT_NTA);
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 48), e);
    }
    _mm_sfence();                   // the stores are globally visible before the chunk is handed to the copy engine / the caller
    if (bytes > body) memcpy(d + body, s + body, bytes - body);
}
#else
extern "C" void lbk_host_stream_copy(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
#endif
