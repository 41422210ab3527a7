// shared-memory Pstream seam (see UPstream.C in this directory).  Only the contiguous-data read
// is provided — it is what the lduMatrix path uses (processorLduInterface::send/receive,
// Pstream::gather/scatter of contiguous types); the token-stream constructors stop with an error.
#include "UIPstream.H"
#include "PstreamBuffers.H"
#include "error.H"
#include "shmWorld.H"

namespace lduShm
{
    void postRecv(int from, char* buf, std::size_t n, int tag);
}

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
