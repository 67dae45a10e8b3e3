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
    ~Mailbox();
    const void *rows() const { return base + MB_HEADER_BYTES; }
    const unsigned long long *count_dev() const { return &reinterpret_cast<const MailboxHeader *>(base)->count; }
};

}  // namespace nbr
