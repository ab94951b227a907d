// Streaming .gz writer for device-resident data: the I/O side of the resampler.
//
// Replaces the mcpl_add_particle loop + mcpl_closeandgzip_outfile of
// clib/src/resamplemcpl.c:40-46.  Device buffers are compressed on the GPU
// (gzip.cu), copied to one of two pinned host buffers and written by a worker thread
// (parallel pwrite of slices) while the caller already produces the next batch, so the
// generator, the PCIe copy and the file system work concurrently.  Small host-side
// pieces (the MCPL header) go through zlib as a gzip member of their own.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

struct GzJob {
  int slot;
  uint64_t bytes, file_off;
  cudaEvent_t ev;
};

}  // namespace kdsb

struct kdsb200_gz_writer_s {
  int fd = -1;
  int device = 0;
  cudaStream_t st = nullptr;
  uint64_t cap = 0, file_off = 0;
  uint32_t member = 0, rec_bytes = 0, tail_bytes = 0;
  bool have_table = false;
  kdsb200_gz_table table;
  unsigned char* d_out[2] = {nullptr, nullptr};
  unsigned char* h_out[2] = {nullptr, nullptr};
  unsigned char* d_scratch = nullptr;
  uint64_t* d_total = nullptr;
  uint64_t* d_hist = nullptr;
  uint64_t* h_small = nullptr;  // pinned: [0] total, [1..258] histogram
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int next_slot = 0, write_threads = 8;
  bool use_mmap = true;
  // worker
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<kdsb::GzJob> queue;
  bool busy[2] = {false, false};
  bool stop = false, failed = false;
  std::string err;
};

namespace kdsb {

static bool pwrite_all(int fd, const unsigned char* p, uint64_t n, uint64_t off) {
  while (n) {
    const ssize_t w = pwrite(fd, p, n > (1u << 30) ? (1u << 30) : n, (off_t)off);
    if (w <= 0) return false;
    p += w;
    off += (uint64_t)w;
    n -= (uint64_t)w;
  }
  return true;
}

// Copy a finished batch into the file through a shared mapping: the file is extended to
// its new length and the slices are memcpy'd by several threads.  Unlike pwrite(), which
// holds the inode lock for the whole copy (measured on tmpfs: 2.6 GB/s whatever the number
// of threads), page faults on a mapping are served in parallel.  Returns false when the
// file system does not support it (the caller falls back to pwrite).
static bool mmap_write(int fd, const unsigned char* src, uint64_t n, uint64_t off, int nthreads) {
  const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
  struct stat sb;  // extend only: a later host member may already lie behind this batch
  if (fstat(fd, &sb) != 0) return false;
  if ((uint64_t)sb.st_size < off + n && ftruncate(fd, (off_t)(off + n)) != 0) return false;
  const uint64_t a0 = off / page * page;
  const size_t len = (size_t)(off + n - a0);
  void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)a0);
  if (m == MAP_FAILED) return false;
  unsigned char* dst = (unsigned char*)m + (off - a0);
  const uint64_t per = ((n + nthreads - 1) / nthreads + page - 1) / page * page;
  std::vector<std::thread> ths;
  for (int i = 0; i < nthreads; i++) {
    const uint64_t lo = std::min<uint64_t>(n, per * i), hi = std::min<uint64_t>(n, per * (i + 1));
    if (hi > lo) ths.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
  }
  for (auto& t : ths) t.join();
  return munmap(m, len) == 0;
}

static void writer_loop(kdsb200_gz_writer_s* w) {
  cudaSetDevice(w->device);  // a new thread starts on device 0; the events live on w->device
  for (;;) {
    GzJob job;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->stop || !w->queue.empty(); });
      if (w->queue.empty()) return;
      job = w->queue.front();
      w->queue.erase(w->queue.begin());
    }
    bool ok = cudaEventSynchronize(job.ev) == cudaSuccess;
    if (ok && w->use_mmap && job.bytes > (8u << 20)) {
      if (!mmap_write(w->fd, w->h_out[job.slot], job.bytes, job.file_off, w->write_threads))
        w->use_mmap = false;  // not supported here: pwrite from now on (this batch included)
      else
        goto written;
    }
    if (ok) {
      const int nt = job.bytes > (8u << 20) ? w->write_threads : 1;
      const uint64_t per = ((job.bytes + nt - 1) / nt + 4095) / 4096 * 4096;
      std::vector<std::thread> ths;
      std::vector<char> oks(nt, 1);
      for (int i = 0; i < nt; i++) {
        const uint64_t lo = std::min<uint64_t>(job.bytes, per * i);
        const uint64_t hi = std::min<uint64_t>(job.bytes, per * (i + 1));
        if (hi <= lo) continue;
        ths.emplace_back([=, &oks] {
          oks[i] = pwrite_all(w->fd, w->h_out[job.slot] + lo, hi - lo, job.file_off + lo);
        });
      }
      for (auto& t : ths) t.join();
      for (char c : oks) ok = ok && c;
    }
  written:
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->busy[job.slot] = false;
      if (!ok && !w->failed) {
        w->failed = true;
        w->err = "gz_writer: copy or write failed";
      }
    }
    w->cv.notify_all();
  }
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

kdsb200_gz_writer* kdsb200_gz_writer_open(const char* path, uint64_t max_batch_bytes,
                                          uint32_t rec_bytes, uint32_t tail_bytes, void* stream) {
  if (max_batch_bytes == 0 || (max_batch_bytes & 3)) {
    set_error("gz_writer_open: max_batch_bytes must be a positive multiple of 4");
    return nullptr;
  }
  if (rec_bytes && (kdsb200_gz_member_bytes() % (256 * rec_bytes) != 0 || rec_bytes % 4 ||
                    tail_bytes % 4 || tail_bytes > rec_bytes || tail_bytes > 256)) {
    set_error("gz_writer_open: unsupported record layout (%u, %u)", rec_bytes, tail_bytes);
    return nullptr;
  }
  kdsb200_gz_writer_s* w = new kdsb200_gz_writer_s();
  w->fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);  // read access: shared mapping
  if (w->fd < 0) {
    set_error("gz_writer_open: cannot create %s", path);
    delete w;
    return nullptr;
  }
  w->st = (cudaStream_t)stream;
  cudaGetDevice(&w->device);
  w->member = kdsb200_gz_member_bytes();
  w->rec_bytes = rec_bytes;
  w->tail_bytes = tail_bytes;
  // Huffman coding never needs more than 15 bits per byte, but a table built from the
  // first batch keeps the typical output below the input; leave generous headroom
  w->cap = (max_batch_bytes + max_batch_bytes / 4 + (1u << 20) + 3) / 4 * 4;
  const uint64_t scratch = kdsb200_gz_scratch_bytes(max_batch_bytes, w->member);
  bool ok = true;
  for (int s = 0; s < 2 && ok; s++) {
    ok = ok && cudaMalloc(&w->d_out[s], w->cap) == cudaSuccess;
    ok = ok && cudaMallocHost(&w->h_out[s], w->cap) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&w->ev[s], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&w->d_scratch, scratch) == cudaSuccess;
  ok = ok && cudaMalloc(&w->d_total, 8) == cudaSuccess;
  ok = ok && cudaMalloc(&w->d_hist, 258 * 8) == cudaSuccess;
  ok = ok && cudaMallocHost(&w->h_small, 259 * 8) == cudaSuccess;
  if (!ok) {
    set_error("gz_writer_open: out of device or pinned memory (batch of %llu bytes)",
              (unsigned long long)max_batch_bytes);
    uint64_t dummy;
    kdsb200_gz_writer_close(w, &dummy);
    return nullptr;
  }
  // batches are copied into the file by KDSB200_WRITE_THREADS threads (default 8) through a
  // shared mapping: 22.5 GB in 3.1 s on tmpfs (4 threads: 4.4 s, 16: no further gain); with
  // pwrite() 2, 8 and 16 threads all gave 2.6 GB/s (one file, one inode lock)
  const char* e = getenv("KDSB200_WRITE_THREADS");
  const int nt = e ? atoi(e) : 8;
  w->write_threads = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
  const char* mode = getenv("KDSB200_WRITE_MODE");  // "pwrite" disables the mapped copy
  w->use_mmap = !(mode && strcmp(mode, "pwrite") == 0);
  w->worker = std::thread(writer_loop, w);
  return w;
}

static int writer_check(kdsb200_gz_writer* w) {
  std::lock_guard<std::mutex> lk(w->mu);
  if (w->failed) {
    set_error("%s", w->err.c_str());
    return 1;
  }
  return 0;
}

int kdsb200_gz_writer_put_host(kdsb200_gz_writer* w, const void* h_data, uint64_t nbytes) {
  if (writer_check(w)) return 1;
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, 6, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
    set_error("gz_writer_put_host: deflateInit2 failed");
    return 1;
  }
  // the same "KD" FEXTRA subfield as the device members carry: the member's own size,
  // patched in once it is known (the header is not covered by the CRC)
  unsigned char extra[8] = {'K', 'D', 4, 0, 0, 0, 0, 0};
  gz_header gh;
  memset(&gh, 0, sizeof(gh));
  gh.os = 255;
  gh.extra = extra;
  gh.extra_len = 8;
  deflateSetHeader(&zs, &gh);
  std::vector<unsigned char> buf(deflateBound(&zs, (uLong)nbytes) + 64);
  zs.next_in = (Bytef*)h_data;
  zs.avail_in = (uInt)nbytes;
  zs.next_out = buf.data();
  zs.avail_out = (uInt)buf.size();
  const int rc = deflate(&zs, Z_FINISH);
  const uint64_t n = zs.total_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) {
    set_error("gz_writer_put_host: deflate failed (%d)", rc);
    return 1;
  }
  if (n >= (1ull << 32) || n < 20 || buf[12] != 'K' || buf[13] != 'D') {
    set_error("gz_writer_put_host: unexpected member layout");
    return 1;
  }
  for (int c = 0; c < 4; c++) buf[16 + c] = (unsigned char)(n >> (8 * c));
  // ordered with the device batches by the running file offset
  if (!pwrite_all(w->fd, buf.data(), n, w->file_off)) {
    set_error("gz_writer_put_host: write failed");
    return 1;
  }
  w->file_off += n;
  return 0;
}

int kdsb200_gz_writer_put_device(kdsb200_gz_writer* w, const void* d_data, uint64_t nbytes) {
  if (writer_check(w)) return 1;
  if (nbytes == 0) return 0;
  const uint64_t max_in = (w->cap - (1u << 20)) / 5 * 4;
  if (nbytes > max_in || (nbytes & 3)) {
    set_error("gz_writer_put_device: batch of %llu bytes (limit %llu, multiple of 4)",
              (unsigned long long)nbytes, (unsigned long long)max_in);
    return 1;
  }
  if (w->rec_bytes && nbytes % w->rec_bytes) {
    set_error("gz_writer_put_device: %llu bytes are not whole %u-byte records",
              (unsigned long long)nbytes, w->rec_bytes);
    return 1;
  }
  if (!w->have_table) {  // one Huffman table for the whole file, from the first batch
    uint64_t sample = nbytes < (64u << 20) ? nbytes : (64u << 20);
    if (w->rec_bytes) sample -= sample % w->rec_bytes;
    if (kdsb200_gz_histogram(d_data, sample, w->member, w->rec_bytes, w->tail_bytes, w->d_hist,
                             w->st))
      return 1;
    KDSB_CUDA(cudaMemcpyAsync(w->h_small + 1, w->d_hist, 258 * 8, cudaMemcpyDeviceToHost, w->st));
    KDSB_CUDA(cudaStreamSynchronize(w->st));
    if (kdsb200_gz_build_table(w->h_small + 1, w->rec_bytes, w->tail_bytes, &w->table)) return 1;
    w->have_table = true;
  }
  const int slot = w->next_slot;
  w->next_slot ^= 1;
  {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return !w->busy[slot] || w->failed; });
    if (w->failed) {
      set_error("%s", w->err.c_str());
      return 1;
    }
  }
  uint64_t total = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    if (kdsb200_gz_compress(d_data, nbytes, w->member, &w->table, w->d_scratch, w->d_out[slot],
                            w->cap, w->d_total, w->st))
      return 1;
    KDSB_CUDA(cudaMemcpyAsync(w->h_small, w->d_total, 8, cudaMemcpyDeviceToHost, w->st));
    KDSB_CUDA(cudaStreamSynchronize(w->st));
    total = w->h_small[0];
    if (total <= w->cap) break;
    if (attempt == 1) {
      set_error("gz_writer_put_device: compressed batch (%llu bytes) exceeds the buffer (%llu)",
                (unsigned long long)total, (unsigned long long)w->cap);
      return 1;
    }
    // this batch does not look like the data the table was built from: rebuild the table
    // from the batch itself (an order-0 code with every symbol present cannot expand its
    // own data by more than the block headers) and encode again
    if (kdsb200_gz_histogram(d_data, nbytes, w->member, w->rec_bytes, w->tail_bytes, w->d_hist,
                             w->st))
      return 1;
    KDSB_CUDA(cudaMemcpyAsync(w->h_small + 1, w->d_hist, 258 * 8, cudaMemcpyDeviceToHost, w->st));
    KDSB_CUDA(cudaStreamSynchronize(w->st));
    if (kdsb200_gz_build_table(w->h_small + 1, w->rec_bytes, w->tail_bytes, &w->table)) return 1;
  }
  KDSB_CUDA(cudaMemcpyAsync(w->h_out[slot], w->d_out[slot], total, cudaMemcpyDeviceToHost, w->st));
  KDSB_CUDA(cudaEventRecord(w->ev[slot], w->st));
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->busy[slot] = true;
    w->queue.push_back({slot, total, w->file_off, w->ev[slot]});
  }
  w->cv.notify_all();
  w->file_off += total;
  return 0;
}

int kdsb200_gz_writer_close(kdsb200_gz_writer* w, uint64_t* bytes_written) {
  if (!w) return 0;
  if (w->worker.joinable()) {
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return (w->queue.empty() && !w->busy[0] && !w->busy[1]) || w->failed; });
      w->stop = true;
    }
    w->cv.notify_all();
    w->worker.join();
  }
  int rc = 0;
  if (w->failed) {
    set_error("%s", w->err.c_str());
    rc = 1;
  }
  if (bytes_written) *bytes_written = w->file_off;
  if (w->fd >= 0 && close(w->fd) != 0 && !rc) {
    set_error("gz_writer_close: close failed");
    rc = 1;
  }
  for (int s = 0; s < 2; s++) {
    if (w->d_out[s]) cudaFree(w->d_out[s]);
    if (w->h_out[s]) cudaFreeHost(w->h_out[s]);
    if (w->ev[s]) cudaEventDestroy(w->ev[s]);
  }
  if (w->d_scratch) cudaFree(w->d_scratch);
  if (w->d_total) cudaFree(w->d_total);
  if (w->d_hist) cudaFree(w->d_hist);
  if (w->h_small) cudaFreeHost(w->h_small);
  delete w;
  return rc;
}

}  // extern "C"
