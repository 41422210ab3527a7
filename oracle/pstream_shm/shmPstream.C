// Test infrastructure (oracle/): a shared-memory implementation of the reference's Pstream seam —
// exactly the symbols of src/Pstream/dummy/{UPstream,UIPread,UOPwrite}.C, what src/Pstream/mpi
// implements over MPI (UPstream.C:64-382, UIPread.C:184-335, UOPwrite.C:37-139) — so that the
// UNMODIFIED reference solvers run as N processes on one host without MPI.
// Rank, size and the path of the mapped file come from the environment (set by oracle/oracle.py).
// Scalar reductions sum in rank order, so every rank gets the bit-identical result.
#include "UPstream.H"
#include "UIPstream.H"
#include "UOPstream.H"
#include "PstreamBuffers.H"
#include "error.H"
#include "PstreamReduceOps.H"
#include "shmWorld.H"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

namespace lduShm
{

static World* w_ = 0;
static int rank_ = 0, size_ = 1;
static int localSense_ = 0;

World* world() { return w_; }
int myRank() { return rank_; }
int nRanks() { return size_; }

static inline void pauseSpin() { sched_yield(); }

// LDU_PSTREAM_STATS=1: seconds this rank waited in barriers / receives, printed at exit
static double waitBarrier_ = 0, waitRecv_ = 0;
static long nBarrier_ = 0, nRecv_ = 0;
static inline double now()
{
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9*t.tv_nsec;
}
struct StatsAtExit
{
    ~StatsAtExit()
    {
        if (getenv("LDU_PSTREAM_STATS"))
        {
            fprintf(stderr, "shm Pstream rank %d: %ld barriers %.3f s, %ld receives %.3f s\n",
                    rank_, nBarrier_, waitBarrier_, nRecv_, waitRecv_);
        }
    }
};
static StatsAtExit statsAtExit_;

void attach()
{
    const char* name = getenv("LDU_PSTREAM_SHM");
    const char* r = getenv("LDU_PSTREAM_RANK");
    const char* s = getenv("LDU_PSTREAM_SIZE");
    if (!name || !r || !s)
    {
        fprintf(stderr, "shm Pstream: LDU_PSTREAM_SHM/RANK/SIZE not set\n");
        abort();
    }
    rank_ = atoi(r);
    size_ = atoi(s);
    int fd = open(name, O_RDWR);   // a zero-filled file of sizeof(World) bytes, made by the launcher
    if (fd < 0) { perror(name); abort(); }
    void* p = mmap(0, sizeof(World), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { perror("mmap"); abort(); }
    w_ = static_cast<World*>(p);
    __sync_fetch_and_add(&w_->arrived, 1);
    while (w_->arrived < size_) pauseSpin();
}

void detach()
{
    if (w_) munmap(w_, sizeof(World));
    w_ = 0;
}

void barrier()
{
    const double t0 = now();
    nBarrier_++;
    localSense_ ^= 1;
    if (__sync_add_and_fetch(&w_->barCount, 1) == size_)
    {
        w_->barCount = 0;
        __sync_synchronize();
        w_->barSense = localSense_;
    }
    else
    {
        while (w_->barSense != localSense_) pauseSpin();
    }
    __sync_synchronize();
    waitBarrier_ += now() - t0;
}

static void ringWrite(Ring& q, const char* src, std::size_t n)
{
    std::size_t done = 0;
    while (done < n)
    {
        const unsigned long long head = q.head, tail = q.tail;
        const std::size_t space = ringBytes - (std::size_t)(head - tail);
        if (!space) { pauseSpin(); continue; }
        std::size_t chunk = n - done < space ? n - done : space;
        const std::size_t off = (std::size_t)(head % ringBytes);
        if (chunk > ringBytes - off) chunk = ringBytes - off;
        memcpy(q.data + off, src + done, chunk);
        __sync_synchronize();
        q.head = head + chunk;
        done += chunk;
    }
}

static void ringRead(Ring& q, char* dst, std::size_t n)
{
    std::size_t done = 0;
    while (done < n)
    {
        const unsigned long long head = q.head, tail = q.tail;
        const std::size_t avail = (std::size_t)(head - tail);
        if (!avail) { pauseSpin(); continue; }
        std::size_t chunk = n - done < avail ? n - done : avail;
        const std::size_t off = (std::size_t)(tail % ringBytes);
        if (chunk > ringBytes - off) chunk = ringBytes - off;
        __sync_synchronize();
        memcpy(dst + done, q.data + off, chunk);
        __sync_synchronize();
        q.tail = tail + chunk;
        done += chunk;
    }
}

void sendBytes(int to, const char* buf, std::size_t n, int tag)
{
    Ring& q = w_->ring[rank_][to];
    int hdr[2] = {tag, (int)n};
    ringWrite(q, reinterpret_cast<const char*>(hdr), sizeof(hdr));
    ringWrite(q, buf, n);
}

bool tryRecvBytes(int from, char* buf, std::size_t n, int tag)
{
    Ring& q = w_->ring[from][rank_];
    if (q.head == q.tail) return false;
    recvBytes(from, buf, n, tag);
    return true;
}

void recvBytes(int from, char* buf, std::size_t n, int tag)
{
    Ring& q = w_->ring[from][rank_];
    int hdr[2];
    const double t0 = now();
    nRecv_++;
    ringRead(q, reinterpret_cast<char*>(hdr), sizeof(hdr));
    waitRecv_ += now() - t0;
    if (hdr[0] != tag || (std::size_t)hdr[1] > n)
    {
        fprintf(stderr, "shm Pstream: rank %d expected tag %d size <= %zu from %d, got tag %d size %d\n",
                rank_, tag, n, from, hdr[0], hdr[1]);
        abort();
    }
    ringRead(q, buf, (std::size_t)hdr[1]);
}

// pending non-blocking receives (sends are buffered and complete at once)
struct Request { int from; char* buf; std::size_t n; int tag; bool done; };
static std::vector<Request> requests_;

void postRecv(int from, char* buf, std::size_t n, int tag)
{
    Request r = {from, buf, n, tag, false};
    requests_.push_back(r);
}

static void complete(Request& r)
{
    if (!r.done)
    {
        recvBytes(r.from, r.buf, r.n, r.tag);
        r.done = true;
    }
}

}   // namespace lduShm


// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

void Foam::UPstream::addValidParOptions(HashTable<string>& validParOptions)
{}


bool Foam::UPstream::init(int& argc, char**& argv)
{
    lduShm::attach();
    myProcNo_ = lduShm::myRank();
    procIDs_.setSize(lduShm::nRanks());
    forAll(procIDs_, procNo)
    {
        procIDs_[procNo] = procNo;
    }
    setParRun();
    initCommunicationSchedule();
    return true;
}


void Foam::UPstream::exit(int errnum)
{
    lduShm::barrier();
    lduShm::detach();
    ::exit(errnum);
}


void Foam::UPstream::abort()
{
    ::abort();
}


static double sumInRankOrder(double v, int k)
{
    lduShm::World* w = lduShm::world();
    w->slot[lduShm::myRank()][k] = v;
    lduShm::barrier();
    double s = 0;
    for (int r = 0; r < lduShm::nRanks(); r++) s += w->slot[r][k];
    lduShm::barrier();
    return s;
}


void Foam::reduce(scalar& Value, const sumOp<scalar>& bop, const int tag)
{
    if (!UPstream::parRun()) return;
    Value = sumInRankOrder(Value, 0);
}


void Foam::reduce(scalar& Value, const minOp<scalar>& bop, const int tag)
{
    if (!UPstream::parRun()) return;
    lduShm::World* w = lduShm::world();
    w->slot[lduShm::myRank()][0] = Value;
    lduShm::barrier();
    scalar m = w->slot[0][0];
    for (int r = 1; r < lduShm::nRanks(); r++) m = min(m, scalar(w->slot[r][0]));
    lduShm::barrier();
    Value = m;
}


void Foam::reduce(vector2D& Value, const sumOp<vector2D>& bop, const int tag)
{
    if (!UPstream::parRun()) return;
    lduShm::World* w = lduShm::world();
    w->slot[lduShm::myRank()][0] = Value.x();
    w->slot[lduShm::myRank()][1] = Value.y();
    lduShm::barrier();
    scalar sx = 0, sy = 0;
    for (int r = 0; r < lduShm::nRanks(); r++)
    {
        sx += w->slot[r][0];
        sy += w->slot[r][1];
    }
    lduShm::barrier();
    Value = vector2D(sx, sy);
}


void Foam::sumReduce(scalar& Value, label& Count, const int tag)
{
    if (!UPstream::parRun()) return;
    lduShm::World* w = lduShm::world();
    w->slot[lduShm::myRank()][0] = Value;
    w->islot[lduShm::myRank()] = Count;
    lduShm::barrier();
    scalar s = 0;
    long long c = 0;
    for (int r = 0; r < lduShm::nRanks(); r++)
    {
        s += w->slot[r][0];
        c += w->islot[r];
    }
    lduShm::barrier();
    Value = s;
    Count = label(c);
}


void Foam::reduce(scalar& Value, const sumOp<scalar>& bop, const int tag, label& request)
{
    reduce(Value, bop, tag);
    request = -1;
}


Foam::label Foam::UPstream::nRequests()
{
    return lduShm::requests_.size();
}


void Foam::UPstream::resetRequests(const label i)
{
    if (i < label(lduShm::requests_.size())) lduShm::requests_.resize(i);
}


void Foam::UPstream::waitRequests(const label start)
{
    for (std::size_t i = start; i < lduShm::requests_.size(); i++) lduShm::complete(lduShm::requests_[i]);
    if (start < label(lduShm::requests_.size())) lduShm::requests_.resize(start);
}


void Foam::UPstream::waitRequest(const label i)
{
    if (i >= 0 && i < label(lduShm::requests_.size())) lduShm::complete(lduShm::requests_[i]);
}


bool Foam::UPstream::finishedRequest(const label i)
{
    if (i < 0 || i >= label(lduShm::requests_.size())) return true;
    lduShm::Request& r = lduShm::requests_[i];
    if (r.done) return true;
    // FIFO per pair: only the oldest pending receive from that rank can complete
    for (label j = 0; j < i; j++)
        if (!lduShm::requests_[j].done && lduShm::requests_[j].from == r.from) return false;
    if (lduShm::tryRecvBytes(r.from, r.buf, r.n, r.tag)) r.done = true;
    return r.done;
}


// ---------------------------------------------------------------------------
// UOPstream::write (what src/Pstream/mpi/UOPwrite.C does with MPI_Bsend/Send/Isend): every write is
// buffered in the ring of the (this rank -> toProcNo) pair and complete on return, whatever the
// commsType
// ---------------------------------------------------------------------------
bool Foam::UOPstream::write
(
    const commsTypes commsType,
    const int toProcNo,
    const char* buf,
    const std::streamsize bufSize,
    const int tag
)
{
    lduShm::sendBytes(toProcNo, buf, bufSize, tag);
    return true;
}


// ---------------------------------------------------------------------------
// UIPstream (src/Pstream/mpi/UIPread.C).  Only the contiguous-data read is provided — it is what the
// lduMatrix path uses (processorLduInterface::send/receive, Pstream::gather/scatter of contiguous
// types); the token-stream constructors stop with an error.
// ---------------------------------------------------------------------------
static void noTokenStreams()
{
    FatalErrorIn("UIPstream::UIPstream")
        << "the shared-memory Pstream of the lduMatrix test harness carries contiguous data only"
        << Foam::abort(Foam::FatalError);
}

Foam::UIPstream::UIPstream
(
    const commsTypes commsType,
    const int fromProcNo,
    DynamicList<char>& externalBuf,
    label& externalBufPosition,
    const int tag,
    const bool clearAtEnd,
    streamFormat format,
    versionNumber version
)
:
    UPstream(commsType),
    Istream(format, version),
    fromProcNo_(fromProcNo),
    externalBuf_(externalBuf),
    externalBufPosition_(externalBufPosition),
    tag_(tag),
    clearAtEnd_(clearAtEnd),
    messageSize_(0)
{
    noTokenStreams();
}

Foam::UIPstream::UIPstream(const int fromProcNo, PstreamBuffers& buffers)
:
    UPstream(buffers.commsType_),
    Istream(buffers.format_, buffers.version_),
    fromProcNo_(fromProcNo),
    externalBuf_(buffers.recvBuf_[fromProcNo]),
    externalBufPosition_(buffers.recvBufPos_[fromProcNo]),
    tag_(buffers.tag_),
    clearAtEnd_(true),
    messageSize_(0)
{
    noTokenStreams();
}

Foam::label Foam::UIPstream::read
(
    const commsTypes commsType,
    const int fromProcNo,
    char* buf,
    const std::streamsize bufSize,
    const int tag
)
{
    if (commsType == nonBlocking)
    {
        lduShm::postRecv(fromProcNo, buf, bufSize, tag);   // completed by UPstream::waitRequests
    }
    else
    {
        lduShm::recvBytes(fromProcNo, buf, bufSize, tag);
    }
    return bufSize;
}
