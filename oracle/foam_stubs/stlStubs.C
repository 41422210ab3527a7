// TEST INFRASTRUCTURE (oracle/build_ref_fv.py).  The reference's two ASCII-STL readers are flex
// sources (surfMesh/surfaceFormats/stl/STLsurfaceFormatASCII.L:395, triSurface/triSurface/
// interfaces/STL/readSTLASCII.L:379); this image has no flex, so libsurfMesh / libtriSurface get
// these two entry points instead.  They fail loudly: no case in tests/ reads an ASCII STL.
#ifdef STUB_surfMesh
#include "STLsurfaceFormatCore.H"
#include "error.H"

bool Foam::fileFormats::STLsurfaceFormatCore::readASCII(istream&, const off_t)
{
    FatalErrorIn("fileFormats::STLsurfaceFormatCore::readASCII(istream&, const off_t)")
        << "ASCII STL reading is not available in this build of the reference (no flex)"
        << exit(FatalError);
    return false;
}
#endif

#ifdef STUB_triSurface
#include "triSurface.H"
#include "error.H"

bool Foam::triSurface::readSTLASCII(const fileName& STLfileName)
{
    FatalErrorIn("triSurface::readSTLASCII(const fileName&)")
        << "ASCII STL reading is not available in this build of the reference (no flex): "
        << STLfileName << exit(FatalError);
    return false;
}
#endif
