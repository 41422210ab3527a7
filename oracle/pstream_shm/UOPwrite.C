// shared-memory Pstream seam (see UPstream.C in this directory): every write is buffered in the
// ring of the (this rank -> toProcNo) pair and complete on return, whatever the commsType
#include "UOPstream.H"
#include "shmWorld.H"

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
