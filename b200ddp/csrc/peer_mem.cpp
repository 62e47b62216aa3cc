#include "peer_mem.h"

#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>
#include <poll.h>
#include <cerrno>
#include <cstring>

#include "common.h"
#include "drv.h"

namespace b200 {

namespace {
size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

CUmemAllocationProp arena_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
}  // namespace

PeerArena::PeerArena(int rank, int world, int device, size_t bytes, const std::string& uid)
    : rank_(rank), world_(world), device_(device), uid_(uid) {
  if (world < 1 || world > kMaxRanks) throw std::runtime_error("PeerArena: world must be in [1, 8]");
  B200_CUDA_CHECK(cudaSetDevice(device));
  B200_CUDA_CHECK(cudaFree(nullptr));  // make sure the primary context exists
  auto& drv = Driver::get();
  CUmemAllocationProp prop = arena_prop(device);
  size_t gran = 0;
  B200_DRV_CHECK(drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  gran_ = gran;
  int mc = 0;
  CUdevice dev;
  B200_DRV_CHECK(drv.DeviceGet(&dev, device));
  if (drv.MulticastCreate && world > 1 &&
      drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && mc) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = world;
    mp.size = round_up(bytes, gran_);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mgran = 0;
    if (drv.MulticastGetGranularity(&mgran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mgran) {
      gran_ = gran_ > mgran ? gran_ : mgran;
      mc_supported_ = true;
    }
  }
  bytes_ = round_up(bytes < 2 * kSignalBytes ? 2 * kSignalBytes : bytes, gran_);
  stride_ = bytes_;
  B200_CUDA_CHECK(cudaHostAlloc((void**)&err_host_, sizeof(int), cudaHostAllocMapped));
  *err_host_ = 0;
  B200_CUDA_CHECK(cudaHostGetDevicePointer((void**)&err_dev_, err_host_, 0));
}

PeerArena::~PeerArena() {
  try { close(); } catch (...) {}
}

std::string PeerArena::sock_name(int r) const { return "b200ddp-" + uid_ + "-" + std::to_string(r); }

void PeerArena::bind_socket() {
  if (world_ == 1 || sock_ >= 0) return;
  sock_ = ::socket(AF_UNIX, SOCK_DGRAM, 0);
  if (sock_ < 0) throw std::runtime_error("PeerArena: socket() failed");
  sockaddr_un addr;
  memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  std::string name = sock_name(rank_);
  // abstract namespace: sun_path[0] == 0, no filesystem entry to clean up
  memcpy(addr.sun_path + 1, name.data(), name.size());
  socklen_t len = offsetof(sockaddr_un, sun_path) + 1 + name.size();
  if (::bind(sock_, (sockaddr*)&addr, len) != 0)
    throw std::runtime_error(std::string("PeerArena: bind failed: ") + strerror(errno));
}

void PeerArena::send_fd(int to_rank, int fd, int tag) {
  sockaddr_un addr;
  memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  std::string name = sock_name(to_rank);
  memcpy(addr.sun_path + 1, name.data(), name.size());
  socklen_t alen = offsetof(sockaddr_un, sun_path) + 1 + name.size();
  int payload[2] = {rank_, tag};
  iovec iov{payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_name = &addr;
  msg.msg_namelen = alen;
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  for (int attempt = 0; attempt < 200; ++attempt) {
    if (::sendmsg(sock_, &msg, 0) >= 0) return;
    if (errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN) break;
    usleep(50 * 1000);  // peer not bound yet
  }
  throw std::runtime_error(std::string("PeerArena: sendmsg failed: ") + strerror(errno));
}

int PeerArena::recv_fd(int* from_rank, int* tag) {
  pollfd pfd{sock_, POLLIN, 0};
  int pr = ::poll(&pfd, 1, 120 * 1000);
  if (pr <= 0) throw std::runtime_error("PeerArena: timed out waiting for a peer's memory handle");
  int payload[2] = {-1, -1};
  iovec iov{payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (::recvmsg(sock_, &msg, 0) < 0) throw std::runtime_error("PeerArena: recvmsg failed");
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_type != SCM_RIGHTS) throw std::runtime_error("PeerArena: message without a file descriptor");
  int fd = -1;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  *from_rank = payload[0];
  *tag = payload[1];
  return fd;
}

void PeerArena::exchange() {
  if (mapped_) return;
  auto& drv = Driver::get();
  CUmemAllocationProp prop = arena_prop(device_);
  CUmemGenericAllocationHandle h;
  B200_DRV_CHECK(drv.MemCreate(&h, bytes_, &prop, 0));
  local_handle_ = h;
  peer_handles_.assign(world_, 0);
  peer_handles_[rank_] = h;

  if (world_ > 1) {
    int fd = -1;
    B200_DRV_CHECK(drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    try {
      for (int r = 0; r < world_; ++r)
        if (r != rank_) send_fd(r, fd, /*tag=*/0);
      for (int got = 0; got < world_ - 1; ++got) {
        int from = -1, tag = -1;
        int pfd = recv_fd(&from, &tag);
        if (from < 0 || from >= world_ || from == rank_ || tag != 0 || peer_handles_[from] != 0) {
          ::close(pfd);
          throw std::runtime_error("PeerArena: unexpected handle message");
        }
        CUmemGenericAllocationHandle ph;
        const CUresult imported = drv.MemImportFromShareableHandle(&ph, (void*)(uintptr_t)pfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        ::close(pfd);                       // the driver keeps its own reference; never leak the descriptor
        B200_DRV_CHECK(imported);
        peer_handles_[from] = ph;
      }
    } catch (...) {
      ::close(fd);
      throw;
    }
    ::close(fd);
  }

  CUdeviceptr va = 0;
  B200_DRV_CHECK(drv.MemAddressReserve(&va, stride_ * world_, gran_, 0, 0));
  base_ = reinterpret_cast<char*>(va);
  CUmemAccessDesc access;
  memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = device_;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int r = 0; r < world_; ++r) {
    B200_DRV_CHECK(drv.MemMap(va + (size_t)r * stride_, bytes_, 0, peer_handles_[r], 0));
    B200_DRV_CHECK(drv.MemSetAccess(va + (size_t)r * stride_, bytes_, &access, 1));
  }
  mapped_ = true;
  B200_CUDA_CHECK(cudaMemset(local(), 0, bytes_ < (64u << 20) ? bytes_ : kSignalBytes));
  B200_CUDA_CHECK(cudaDeviceSynchronize());
}

void PeerArena::multicast_create() {
  if (!mc_supported_) return;
  auto& drv = Driver::get();
  if (rank_ == 0) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = world_;
    mp.size = bytes_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle mh;
    B200_DRV_CHECK(drv.MulticastCreate(&mh, &mp));
    mc_handle_ = mh;
    int fd = -1;
    B200_DRV_CHECK(drv.MemExportToShareableHandle(&fd, mh, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    try {
      for (int r = 1; r < world_; ++r) send_fd(r, fd, /*tag=*/1);
    } catch (...) {
      ::close(fd);
      throw;
    }
    ::close(fd);
  } else {
    int from = -1, tag = -1;
    int fd = recv_fd(&from, &tag);
    if (from != 0 || tag != 1) {
      ::close(fd);
      throw std::runtime_error("PeerArena: unexpected multicast message");
    }
    CUmemGenericAllocationHandle mh;
    const CUresult imported = drv.MemImportFromShareableHandle(&mh, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    ::close(fd);
    B200_DRV_CHECK(imported);
    mc_handle_ = mh;
  }
}

void PeerArena::multicast_add_device() {
  if (!mc_supported_ || !mc_handle_) return;
  auto& drv = Driver::get();
  CUdevice dev;
  B200_DRV_CHECK(drv.DeviceGet(&dev, device_));
  B200_DRV_CHECK(drv.MulticastAddDevice(mc_handle_, dev));
}

void PeerArena::multicast_bind() {
  if (!mc_supported_ || !mc_handle_) return;
  auto& drv = Driver::get();
  B200_DRV_CHECK(drv.MulticastBindMem(mc_handle_, 0, local_handle_, 0, bytes_, 0));
  mc_bound_ = true;
  CUdeviceptr va = 0;
  B200_DRV_CHECK(drv.MemAddressReserve(&va, bytes_, gran_, 0, 0));
  B200_DRV_CHECK(drv.MemMap(va, bytes_, 0, mc_handle_, 0));
  CUmemAccessDesc access;
  memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = device_;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  B200_DRV_CHECK(drv.MemSetAccess(va, bytes_, &access, 1));
  mc_base_ = reinterpret_cast<char*>(va);
}

void PeerArena::disable_multicast() {
  auto& drv = Driver::get();
  if (mc_base_) {
    drv.MemUnmap((CUdeviceptr)mc_base_, bytes_);
    drv.MemAddressFree((CUdeviceptr)mc_base_, bytes_);
    mc_base_ = nullptr;
  }
  if (mc_bound_ && drv.MulticastUnbind) {
    CUdevice dev;
    if (drv.DeviceGet(&dev, device_) == CUDA_SUCCESS) drv.MulticastUnbind(mc_handle_, dev, 0, bytes_);
    mc_bound_ = false;
  }
  if (mc_handle_) {
    drv.MemRelease(mc_handle_);
    mc_handle_ = 0;
  }
  mc_supported_ = false;
}

size_t PeerArena::alloc(size_t nbytes, size_t align) {
  size_t off = round_up(bump_, align);
  if (off + nbytes > bytes_)
    throw std::runtime_error("PeerArena: out of symmetric memory (need " + std::to_string(off + nbytes) +
                             " of " + std::to_string(bytes_) + " bytes); set B200DDP_ARENA_MB to a larger arena");
  bump_ = off + nbytes;
  return off;
}

void PeerArena::close() {
  auto& drv = Driver::get();
  if (mapped_) cudaDeviceSynchronize();
  disable_multicast();
  if (mapped_) {
    for (int r = 0; r < world_; ++r) drv.MemUnmap((CUdeviceptr)(base_ + (size_t)r * stride_), bytes_);
    drv.MemAddressFree((CUdeviceptr)base_, stride_ * world_);
    for (int r = 0; r < world_; ++r)
      if (peer_handles_[r]) drv.MemRelease(peer_handles_[r]);
    peer_handles_.clear();
    base_ = nullptr;
    mapped_ = false;
  }
  if (sock_ >= 0) { ::close(sock_); sock_ = -1; }
  if (err_host_) { cudaFreeHost(err_host_); err_host_ = nullptr; err_dev_ = nullptr; }
}

}  // namespace b200
