// mailbox.cuh -- one rank's halo mailbox (multi-GPU tile path, mailbox.cu).
#pragma once
#include "common.cuh"

namespace nbr {

constexpr int MB_MAX_WORLD = 16;
constexpr int MB_HEADER_BYTES = 4096;

// first bytes of every mailbox allocation; the peers write `cursor`, their slots of `halo_done`, `box_epoch`
// and `boxes`, everything else is local
struct MailboxHeader {
    unsigned long long cursor;                       // rows reserved by the pushers of this epoch
    unsigned long long pad0[15];
    unsigned long long halo_done[MB_MAX_WORLD];      // epoch of the last finished push of every peer
    unsigned long long box_epoch[MB_MAX_WORLD];      // epoch of the box in slot r
    double boxes[MB_MAX_WORLD][8];                   // lo[3], hi[3], n_points, spare
    unsigned long long count;                        // rows received this epoch (halo_wait)
    unsigned long long overflow;                     // sticky: rows dropped because the mailbox was full
    unsigned long long timeout;                      // sticky: a wait gave up
    unsigned long long blocks_done;                  // push kernel: blocks finished (the last one signals)
    unsigned long long gather_done[MB_MAX_WORLD];    // epoch of the last finished row gather of every peer (gather buffers)
};
static_assert(sizeof(MailboxHeader) <= MB_HEADER_BYTES, "mailbox header");

struct Mailbox {
    int rank = 0, world = 1, dtype = NBR_F32, device = 0;
    int64_t capacity = 0;                            // rows this rank can receive per step
    int64_t capacity_of[MB_MAX_WORLD] = {0};         // rows every peer can receive (0: same as this rank's)
    unsigned char *base = nullptr;                   // own allocation: header | rows
    unsigned char *peer[MB_MAX_WORLD] = {nullptr};   // every rank's allocation as mapped into this process
    bool opened[MB_MAX_WORLD] = {false};             // mapped through cudaIpcOpenMemHandle
    unsigned long long epoch = 0;
    double *box_dev = nullptr;                       // this rank's bounding box (device, 8 doubles)
    double *host_boxes = nullptr;                    // pinned: [MB_MAX_WORLD][8] boxes, then 8 status words
    // feature all-gather through peer stores: every rank owns a STAGING buffer for the rows of all ranks (rank order,
    // every rank's share in that rank's own processing order) + their row numbers, mapped by the others like the mailbox
    // itself; the feature kernel writes every finished row into all of them, the owner puts them in place afterwards
    unsigned char *gather_base = nullptr;
    size_t gather_bytes = 0;
    unsigned char *gather_peer[MB_MAX_WORLD] = {nullptr};
    size_t gather_peer_bytes[MB_MAX_WORLD] = {0};
    bool gather_opened[MB_MAX_WORLD] = {false};
    void gather_release();
    ~Mailbox();
    const void *rows() const { return base + MB_HEADER_BYTES; }
    const unsigned long long *count_dev() const { return &reinterpret_cast<const MailboxHeader *>(base)->count; }
};

// destinations of finished feature rows (rows3.cu) besides the rank's own result: the i-th row the launch finishes
// (processing order) goes to base[d] + (row_offset + i) * row bytes and its row number to perm[d][row_offset + i], for
// every d != self.  contiguous per warp: scattered 80-byte stores over gigabytes of PEER memory miss the TLB on every
// row (measured: 226 ms instead of 6 for 3 peers x 10M rows), contiguous ones do not
struct RowDests {
    unsigned char *base[MB_MAX_WORLD];
    uint32_t *perm[MB_MAX_WORLD];
    long long row_offset;
    int n;
    int self;
};

int halo_wait(Mailbox *M, cudaStream_t stream);
// signals "my rows of this epoch are in your buffer" to every peer, then waits for every peer's signal
int gather_finish(Mailbox *M, cudaStream_t stream);
// staging layout for `total` rows of row_bytes: rows, then (256-byte aligned) the row numbers
inline size_t gather_perm_offset(int64_t total, size_t row_bytes) { return ((size_t)total * row_bytes + 255) & ~(size_t)255; }
inline size_t gather_bytes_needed(int64_t total, size_t row_bytes) { return gather_perm_offset(total, row_bytes) + (size_t)total * 4; }
// finished rows of this rank (tile order, device) -> every peer's staging buffer, with identity row numbers
int gather_push_rows(const RowDests *D, const void *rows, int64_t n, size_t row_bytes, cudaStream_t stream);
// staged rows of the peers -> their places in out_all (rows of all ranks, rank order, every share in tile order)
int gather_unpermute(const Mailbox *M, const int64_t *row_offsets, size_t row_bytes, void *out_all, cudaStream_t stream);

}  // namespace nbr
