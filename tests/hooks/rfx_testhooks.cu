// rfx_testhooks.cu -- TEST-ONLY access to intermediate device state of a libreflexiv_cuda context (packed reads, super-k-mer
// records): copies to host.  Built as tests/hooks/librfx_testhooks.so from the library's internal header; not part of
// the product library or its ABI.
#include "../../reflexiv_b200/csrc/rfx_internal.h"

using namespace rfx;

int rfx::ctx_fail(Ctx* c, int code, const char* fmt, ...) { (void)fmt; if (c) c->err = "test hook failed"; return code; }

extern "C" {
int rfx_debug_reads(rfx_ctx* c, uint64_t* n_reads, uint64_t* total_words, uint32_t* lens, uint64_t* word_offsets, uint64_t* words) {
    if (!c) return RFX_E_INVALID;
    cudaSetDevice(c->prm.device);
    if (n_reads) *n_reads = c->n_reads;
    if (total_words) *total_words = c->n_words;
    if (lens && c->n_reads) RFX_CUDA(c, cudaMemcpy(lens, c->rd_len.p, c->n_reads * 4, cudaMemcpyDeviceToHost));
    if (word_offsets && c->n_reads) RFX_CUDA(c, cudaMemcpy(word_offsets, c->rd_woff.p, c->n_reads * 8, cudaMemcpyDeviceToHost));
    if (words && c->n_words) RFX_CUDA(c, cudaMemcpy(words, c->packed.p, c->n_words * 8, cudaMemcpyDeviceToHost));
    return RFX_OK;
}

int rfx_debug_records(rfx_ctx* c, uint64_t* n_records, uint32_t* n_bins, uint64_t* bin_offsets, uint64_t* records) {
    if (!c) return RFX_E_INVALID;
    if (!c->have_records) return ctx_fail(c, RFX_E_STATE, "no records");
    cudaSetDevice(c->prm.device);
    if (n_records) *n_records = c->n_records;
    if (n_bins) *n_bins = c->n_bins;
    if (bin_offsets) RFX_CUDA(c, cudaMemcpy(bin_offsets, c->bin_off.p, ((size_t)c->n_bins + 1) * 8, cudaMemcpyDeviceToHost));
    if (records && c->n_records) RFX_CUDA(c, cudaMemcpy(records, c->records.p, c->n_records * c->recw * 8, cudaMemcpyDeviceToHost));
    return RFX_OK;
}

}
