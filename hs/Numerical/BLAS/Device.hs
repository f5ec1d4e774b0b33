{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE BangPatterns #-}
{- | Numerical.BLAS.Device -- the Level-1 routines of "Numerical.BLAS.Single" on an
NVIDIA B200, behind the C ABI of @include/lbk.h@ (@lambda_blas_b200/liblbk.so@).

The call surface is the reference's (src/Numerical/BLAS/Single.hs:41-61): the
polymorphic names (@dot@, @axpy@, ...) with their deprecated aliases for Float
(@sdot@, ..., Single.hs:84,113,129,147,165,201,222,421,441,484) and Double (@ddot@,
..., Single.hs:86,115,131,149,167,203,224,423,443,486): same argument order, same
results -- with @Vector a@ replaced by the device-resident 'DVector' @a@ and the
result in 'IO', because the work is enqueued on a CUDA stream.  Where the reference
is polymorphic over @Num a@/@Storable a@, this module is polymorphic over 'Elt',
the class of the two element types the C ABI is instantiated for.  Like the reference the vector
routines are pure in effect: they return NEW device vectors and leave their
arguments untouched (Util.hs:134-142); the in-place forms of the C ABI are
exported with a trailing underscore.

The marshalling idiom is the one of the reference's own FFI module,
test/Foreign.hs:363-396 (@alloca@ / @peek@ for out-parameters, @withForeignPtr@
for buffers), except that scalars are passed by value and every routine returns
a status code.

NOTE: this image has no GHC / stack / cabal (SURVEY.md, Appendix C), so this
module is delivered as source and has not been compiled here.  Every symbol it
imports is exercised, with the same argument order, by the ctypes binding
lambda_blas_b200/_ffi.py, which is what the test-suite drives.
-}
module Numerical.BLAS.Device
    ( -- * Context and device vectors
      Context, withContext, newContext, closeContext, sync
    , Elt, DVector, dvLength, toDevice, fromDevice, cloneD
      -- * Products and norms
    , dot, sdsdot, asum, nrm2, iamax
    , sdot, ddot, sasum, dasum, snrm2, dnrm2, isamax, idamax
      -- * Pure vector results (new device vectors)
    , scal, copy, swap, axpy, rot, rotm
    , sscal, dscal, scopy, dcopy, sswap, dswap, saxpy, daxpy, srot, drot, srotm, drotm
      -- * In-place forms
    , axpy_, scal_, copy_, swap_, rot_, rotm_
      -- * Host scalars
    , rotg, rotmg, srotg, drotg, srotmg, drotmg
    , LbkError(..)
    ) where

import Control.Exception (Exception, bracket, throwIO)
import Control.Monad (when)
import qualified Data.Vector.Storable as V
import qualified Data.Vector.Storable.Mutable as VM
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types
import Foreign.ForeignPtr
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, withArray)
import Foreign.Ptr
import Foreign.Storable (Storable, peek)
import Numerical.BLAS.Types (GivensRot(..), ModGivensRot(..))

-- ---------------------------------------------------------------------------
-- opaque handles
-- ---------------------------------------------------------------------------
data LbkCtx
data LbkVec a          -- phantom: the element type of the device buffer

-- | One GPU, one stream, reduction scratch (and, after @lbk_comm_init@, a communicator).
newtype Context = Context (Ptr LbkCtx)

-- | A device-resident @Vector a@: a finalised handle plus the context that owns it.
data DVector a = DVector { dvCtx :: !Context, dvPtr :: !(ForeignPtr (LbkVec a)), dvLength :: !Int }

data LbkError = LbkError Int String deriving Show
instance Exception LbkError

-- ---------------------------------------------------------------------------
-- the C ABI (include/lbk.h).  `unsafe` only for calls that merely enqueue or are
-- O(1) on the host; anything that blocks on the GPU is a safe call
-- (test/Foreign.hs binds both flavours; test/Bench.hs:122-126 times both).
-- ---------------------------------------------------------------------------
type V a = Ptr (LbkVec a)

foreign import ccall        "lbk_init"        c_init        :: CInt -> Ptr (Ptr LbkCtx) -> IO CInt
foreign import ccall        "lbk_shutdown"    c_shutdown    :: Ptr LbkCtx -> IO CInt
foreign import ccall        "lbk_sync"        c_sync        :: Ptr LbkCtx -> IO CInt
foreign import ccall unsafe "lbk_last_error"  c_last_error  :: Ptr LbkCtx -> IO CString
foreign import ccall        "&lbk_vec_free"   p_vec_free    :: FunPtr (Ptr () -> IO ())

-- Float instance (s / is prefix)
foreign import ccall        "lbk_vec_alloc"    s_vec_alloc    :: Ptr LbkCtx -> Int -> Ptr (V Float) -> IO CInt
foreign import ccall        "lbk_vec_upload"   s_vec_upload   :: V Float -> Ptr Float -> Int -> IO CInt
foreign import ccall        "lbk_vec_download" s_vec_download :: V Float -> Ptr Float -> Int -> IO CInt
foreign import ccall        "lbk_vec_clone"    s_vec_clone    :: V Float -> Ptr (V Float) -> IO CInt
foreign import ccall "lbk_sdot"   s_dot   :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_sdsdot" c_sdsdot :: Ptr LbkCtx -> Int -> Float -> V Float -> Int -> V Float -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_sasum"  s_asum  :: Ptr LbkCtx -> Int -> V Float -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_snrm2"  s_nrm2  :: Ptr LbkCtx -> Int -> V Float -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_isamax" s_iamax :: Ptr LbkCtx -> Int -> V Float -> Int -> Ptr Int -> IO CInt
foreign import ccall unsafe "lbk_saxpy" s_axpy :: Ptr LbkCtx -> Int -> Float -> V Float -> Int -> V Float -> Int -> IO CInt
foreign import ccall unsafe "lbk_sscal" s_scal :: Ptr LbkCtx -> Int -> Float -> V Float -> Int -> IO CInt
foreign import ccall unsafe "lbk_scopy" s_copy :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> IO CInt
foreign import ccall unsafe "lbk_sswap" s_swap :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> IO CInt
foreign import ccall unsafe "lbk_srot"  s_rot  :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> Float -> Float -> IO CInt
foreign import ccall unsafe "lbk_srotm" s_rotm :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> Ptr Float -> IO CInt
foreign import ccall unsafe "lbk_saxpy_oop" s_axpy_oop :: Ptr LbkCtx -> Int -> Float -> V Float -> Int -> V Float -> Int -> V Float -> IO CInt
foreign import ccall unsafe "lbk_sscal_oop" s_scal_oop :: Ptr LbkCtx -> Int -> Float -> V Float -> Int -> V Float -> IO CInt
foreign import ccall unsafe "lbk_scopy_oop" s_copy_oop :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> V Float -> IO CInt
foreign import ccall unsafe "lbk_sswap_oop" s_swap_oop :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> V Float -> V Float -> IO CInt
foreign import ccall unsafe "lbk_srot_oop"  s_rot_oop  :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> Float -> Float -> V Float -> V Float -> IO CInt
foreign import ccall unsafe "lbk_srotm_oop" s_rotm_oop :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> Ptr Float -> V Float -> V Float -> IO CInt
foreign import ccall unsafe "lbk_srotg"  s_rotg  :: Float -> Float -> Ptr Float -> IO CInt
foreign import ccall unsafe "lbk_srotmg" s_rotmg :: Float -> Float -> Float -> Float -> Ptr Float -> Ptr Float -> IO CInt

-- Double instance (d / id prefix)
foreign import ccall        "lbk_vec_alloc_f64" d_vec_alloc   :: Ptr LbkCtx -> Int -> Ptr (V Double) -> IO CInt
foreign import ccall        "lbk_vec_upload"   d_vec_upload   :: V Double -> Ptr Double -> Int -> IO CInt
foreign import ccall        "lbk_vec_download" d_vec_download :: V Double -> Ptr Double -> Int -> IO CInt
foreign import ccall        "lbk_vec_clone"    d_vec_clone    :: V Double -> Ptr (V Double) -> IO CInt
foreign import ccall "lbk_ddot"   d_dot   :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> Ptr Double -> IO CInt
foreign import ccall "lbk_dasum"  d_asum  :: Ptr LbkCtx -> Int -> V Double -> Int -> Ptr Double -> IO CInt
foreign import ccall "lbk_dnrm2"  d_nrm2  :: Ptr LbkCtx -> Int -> V Double -> Int -> Ptr Double -> IO CInt
foreign import ccall "lbk_idamax" d_iamax :: Ptr LbkCtx -> Int -> V Double -> Int -> Ptr Int -> IO CInt
foreign import ccall unsafe "lbk_daxpy" d_axpy :: Ptr LbkCtx -> Int -> Double -> V Double -> Int -> V Double -> Int -> IO CInt
foreign import ccall unsafe "lbk_dscal" d_scal :: Ptr LbkCtx -> Int -> Double -> V Double -> Int -> IO CInt
foreign import ccall unsafe "lbk_dcopy" d_copy :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> IO CInt
foreign import ccall unsafe "lbk_dswap" d_swap :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> IO CInt
foreign import ccall unsafe "lbk_drot"  d_rot  :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> Double -> Double -> IO CInt
foreign import ccall unsafe "lbk_drotm" d_rotm :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> Ptr Double -> IO CInt
foreign import ccall unsafe "lbk_daxpy_oop" d_axpy_oop :: Ptr LbkCtx -> Int -> Double -> V Double -> Int -> V Double -> Int -> V Double -> IO CInt
foreign import ccall unsafe "lbk_dscal_oop" d_scal_oop :: Ptr LbkCtx -> Int -> Double -> V Double -> Int -> V Double -> IO CInt
foreign import ccall unsafe "lbk_dcopy_oop" d_copy_oop :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> V Double -> IO CInt
foreign import ccall unsafe "lbk_dswap_oop" d_swap_oop :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> V Double -> V Double -> IO CInt
foreign import ccall unsafe "lbk_drot_oop"  d_rot_oop  :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> Double -> Double -> V Double -> V Double -> IO CInt
foreign import ccall unsafe "lbk_drotm_oop" d_rotm_oop :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> Ptr Double -> V Double -> V Double -> IO CInt
foreign import ccall unsafe "lbk_drotg"  d_rotg  :: Double -> Double -> Ptr Double -> IO CInt
foreign import ccall unsafe "lbk_drotmg" d_rotmg :: Double -> Double -> Double -> Double -> Ptr Double -> Ptr Double -> IO CInt

-- | The element types the C ABI is instantiated for: the reference's @Float@ and @Double@
-- instances (Single.hs:83-87).  Each method is one entry point of include/lbk.h.
class (Storable a, RealFrac a) => Elt a where
    cVecAlloc    :: Ptr LbkCtx -> Int -> Ptr (V a) -> IO CInt
    cVecUpload   :: V a -> Ptr a -> Int -> IO CInt
    cVecDownload :: V a -> Ptr a -> Int -> IO CInt
    cVecClone    :: V a -> Ptr (V a) -> IO CInt
    cDot     :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> Ptr a -> IO CInt
    cAsum    :: Ptr LbkCtx -> Int -> V a -> Int -> Ptr a -> IO CInt
    cNrm2    :: Ptr LbkCtx -> Int -> V a -> Int -> Ptr a -> IO CInt
    cIamax   :: Ptr LbkCtx -> Int -> V a -> Int -> Ptr Int -> IO CInt
    cAxpy    :: Ptr LbkCtx -> Int -> a -> V a -> Int -> V a -> Int -> IO CInt
    cScal    :: Ptr LbkCtx -> Int -> a -> V a -> Int -> IO CInt
    cCopy    :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> IO CInt
    cSwap    :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> IO CInt
    cRot     :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> a -> a -> IO CInt
    cRotm    :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> Ptr a -> IO CInt
    cAxpyOop :: Ptr LbkCtx -> Int -> a -> V a -> Int -> V a -> Int -> V a -> IO CInt
    cScalOop :: Ptr LbkCtx -> Int -> a -> V a -> Int -> V a -> IO CInt
    cCopyOop :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> V a -> IO CInt
    cSwapOop :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> V a -> V a -> IO CInt
    cRotOop  :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> a -> a -> V a -> V a -> IO CInt
    cRotmOop :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> Ptr a -> V a -> V a -> IO CInt
    cRotg    :: a -> a -> Ptr a -> IO CInt
    cRotmg   :: a -> a -> a -> a -> Ptr a -> Ptr a -> IO CInt

instance Elt Float where
    cVecAlloc = s_vec_alloc; cVecUpload = s_vec_upload; cVecDownload = s_vec_download; cVecClone = s_vec_clone
    cDot = s_dot; cAsum = s_asum; cNrm2 = s_nrm2; cIamax = s_iamax
    cAxpy = s_axpy; cScal = s_scal; cCopy = s_copy; cSwap = s_swap; cRot = s_rot; cRotm = s_rotm
    cAxpyOop = s_axpy_oop; cScalOop = s_scal_oop; cCopyOop = s_copy_oop; cSwapOop = s_swap_oop
    cRotOop = s_rot_oop; cRotmOop = s_rotm_oop; cRotg = s_rotg; cRotmg = s_rotmg

instance Elt Double where
    cVecAlloc = d_vec_alloc; cVecUpload = d_vec_upload; cVecDownload = d_vec_download; cVecClone = d_vec_clone
    cDot = d_dot; cAsum = d_asum; cNrm2 = d_nrm2; cIamax = d_iamax
    cAxpy = d_axpy; cScal = d_scal; cCopy = d_copy; cSwap = d_swap; cRot = d_rot; cRotm = d_rotm
    cAxpyOop = d_axpy_oop; cScalOop = d_scal_oop; cCopyOop = d_copy_oop; cSwapOop = d_swap_oop
    cRotOop = d_rot_oop; cRotmOop = d_rotm_oop; cRotg = d_rotg; cRotmg = d_rotmg

-- ---------------------------------------------------------------------------
-- helpers
-- ---------------------------------------------------------------------------
check :: Context -> CInt -> IO ()
check (Context c) st = when (st /= 0) $ do
    msg <- c_last_error c >>= peekCString
    throwIO (LbkError (fromIntegral st) msg)

newContext :: Int -> IO Context
newContext dev = alloca $ \ pp -> do
    st <- c_init (fromIntegral dev) pp
    when (st /= 0) $ c_last_error nullPtr >>= peekCString >>= throwIO . LbkError (fromIntegral st)
    Context <$> peek pp

closeContext :: Context -> IO ()
closeContext (Context c) = () <$ c_shutdown c

withContext :: Int -> (Context -> IO a) -> IO a
withContext dev = bracket (newContext dev) closeContext

sync :: Context -> IO ()
sync ctx@(Context c) = c_sync c >>= check ctx

adopt :: Context -> Int -> V a -> IO (DVector a)
adopt ctx n p = do fp <- newForeignPtr (castFunPtr p_vec_free) p
                   return (DVector ctx fp n)

allocD :: Elt a => Context -> Int -> IO (DVector a)
allocD ctx@(Context c) n = alloca $ \ pp -> do
    cVecAlloc c n pp >>= check ctx
    peek pp >>= adopt ctx n

-- | Explicit host -> device transfer (the north star's "explicit host transfer").
toDevice :: Elt a => Context -> V.Vector a -> IO (DVector a)
toDevice ctx v = do
    d <- allocD ctx (V.length v)
    V.unsafeWith v $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        cVecUpload dp hp (V.length v) >>= check ctx
    return d

-- | Explicit device -> host transfer.
fromDevice :: Elt a => DVector a -> IO (V.Vector a)
fromDevice d = do
    mv <- VM.new (dvLength d)
    VM.unsafeWith mv $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        cVecDownload dp hp (dvLength d) >>= check (dvCtx d)
    V.unsafeFreeze mv

cloneD :: Elt a => DVector a -> IO (DVector a)
cloneD d = alloca $ \ pp -> withForeignPtr (dvPtr d) $ \ dp -> do
    cVecClone dp pp >>= check (dvCtx d)
    peek pp >>= adopt (dvCtx d) (dvLength d)

with2 :: DVector a -> DVector a -> (Ptr LbkCtx -> V a -> V a -> IO b) -> IO b
with2 x y f = let Context c = dvCtx x in
    withForeignPtr (dvPtr x) $ \ px -> withForeignPtr (dvPtr y) $ \ py -> f c px py

-- ---------------------------------------------------------------------------
-- reductions (Single.hs:73-87, 155-168, 174-204, 211-225, 231-262)
-- ---------------------------------------------------------------------------
dot :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> IO a
dot n x incx y incy = alloca $ \ po -> with2 x y $ \ c px py -> do
    cDot c n px incx py incy po >>= check (dvCtx x)
    peek po

-- | Float only, like the reference (Single.hs:231-262).
sdsdot :: Int -> Float -> DVector Float -> Int -> DVector Float -> Int -> IO Float
sdsdot n sb x incx y incy = alloca $ \ po -> with2 x y $ \ c px py -> do
    c_sdsdot c n sb px incx py incy po >>= check (dvCtx x)
    peek po

ivi :: Elt a => (Ptr LbkCtx -> Int -> V a -> Int -> Ptr a -> IO CInt) -> Int -> DVector a -> Int -> IO a
ivi f n x incx = let Context c = dvCtx x in
    alloca $ \ po -> withForeignPtr (dvPtr x) $ \ px -> do
        f c n px incx po >>= check (dvCtx x)
        peek po

asum, nrm2 :: Elt a => Int -> DVector a -> Int -> IO a
asum = ivi cAsum
nrm2 = ivi cNrm2

-- | 0-based logical index, -1 when n<1 or incx<1 (Single.hs:217-220).
iamax :: Elt a => Int -> DVector a -> Int -> IO Int
iamax n x incx = let Context c = dvCtx x in
    alloca $ \ po -> withForeignPtr (dvPtr x) $ \ px -> do
        cIamax c n px incx po >>= check (dvCtx x)
        peek po

-- ---------------------------------------------------------------------------
-- pure vector results: new device vectors, arguments untouched (Util.hs:134-142)
-- ---------------------------------------------------------------------------
scal :: Elt a => Int -> a -> DVector a -> Int -> IO (DVector a)
scal n a x incx = do
    o <- allocD (dvCtx x) (dvLength x)
    with2 x o $ \ c px po -> cScalOop c n a px incx po >>= check (dvCtx x)
    return o

copy :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> IO (DVector a)
copy n x incx y incy = do
    o <- allocD (dvCtx y) (dvLength y)
    with2 x y $ \ c px py -> withForeignPtr (dvPtr o) $ \ po ->
        cCopyOop c n px incx py incy po >>= check (dvCtx x)
    return o

axpy :: Elt a => Int -> a -> DVector a -> Int -> DVector a -> Int -> IO (DVector a)
axpy n a x incx y incy = do
    o <- allocD (dvCtx y) (dvLength y)
    with2 x y $ \ c px py -> withForeignPtr (dvPtr o) $ \ po ->
        cAxpyOop c n a px incx py incy po >>= check (dvCtx x)
    return o

pair :: Elt a => DVector a -> DVector a -> (Ptr LbkCtx -> V a -> V a -> V a -> V a -> IO CInt) -> IO (DVector a, DVector a)
pair x y f = do
    xo <- allocD (dvCtx x) (dvLength x)
    yo <- allocD (dvCtx y) (dvLength y)
    with2 x y $ \ c px py -> with2 xo yo $ \ _ pxo pyo -> f c px py pxo pyo >>= check (dvCtx x)
    return (xo, yo)

swap :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> IO (DVector a, DVector a)
swap n x incx y incy = pair x y $ \ c px py pxo pyo -> cSwapOop c n px incx py incy pxo pyo

rot :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> a -> a -> IO (DVector a, DVector a)
rot n x incx y incy cs sn = pair x y $ \ c px py pxo pyo -> cRotOop c n px incx py incy cs sn pxo pyo

-- | BLAS sparam [flag,h11,h21,h12,h22], filled exactly as test/Foreign.hs:276-280 does.
sparam :: Num a => ModGivensRot a -> [a]
sparam FLAGNEG2 = [-2, 1, 0, 0, 1]
sparam m@FLAGNEG1{} = [-1, h11 m, h21 m, h12 m, h22 m]
sparam m@FLAG0{}    = [0, 0, h21 m, h12 m, 0]
sparam m@FLAG1{}    = [1, h11 m, 0, 0, h22 m]

-- | FLAGNEG2 hands back the arguments themselves, as the reference does (Single.hs:473).
rotm :: Elt a => ModGivensRot a -> Int -> DVector a -> Int -> DVector a -> Int -> IO (DVector a, DVector a)
rotm FLAGNEG2 _ x _ y _ = return (x, y)
rotm flags n x incx y incy = withArray (sparam flags) $ \ pp ->
    pair x y $ \ c px py pxo pyo -> cRotmOop c n px incx py incy pp pxo pyo

-- ---------------------------------------------------------------------------
-- in-place forms (enqueue only)
-- ---------------------------------------------------------------------------
axpy_ :: Elt a => Int -> a -> DVector a -> Int -> DVector a -> Int -> IO ()
axpy_ n a x incx y incy = with2 x y $ \ c px py -> cAxpy c n a px incx py incy >>= check (dvCtx x)

scal_ :: Elt a => Int -> a -> DVector a -> Int -> IO ()
scal_ n a x incx = let Context c = dvCtx x in
    withForeignPtr (dvPtr x) $ \ px -> cScal c n a px incx >>= check (dvCtx x)

copy_, swap_ :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> IO ()
copy_ n x incx y incy = with2 x y $ \ c px py -> cCopy c n px incx py incy >>= check (dvCtx x)
swap_ n x incx y incy = with2 x y $ \ c px py -> cSwap c n px incx py incy >>= check (dvCtx x)

rot_ :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> a -> a -> IO ()
rot_ n x incx y incy cs sn = with2 x y $ \ c px py -> cRot c n px incx py incy cs sn >>= check (dvCtx x)

rotm_ :: Elt a => ModGivensRot a -> Int -> DVector a -> Int -> DVector a -> Int -> IO ()
rotm_ flags n x incx y incy = withArray (sparam flags) $ \ pp ->
    with2 x y $ \ c px py -> cRotm c n px incx py incy pp >>= check (dvCtx x)

-- ---------------------------------------------------------------------------
-- host scalars (Single.hs:274-300, 318-357); pure, like the reference's
-- ---------------------------------------------------------------------------
rotg :: Elt a => a -> a -> IO (GivensRot a)
rotg a b = allocaArray 4 $ \ po -> do
    _ <- cRotg a b po
    [r, z, c, s] <- peekArray 4 po
    return (GIVENSROT (r, z, c, s))

rotmg :: Elt a => a -> a -> a -> a -> IO (ModGivensRot a)
rotmg sd1 sd2 sx1 sy1 = allocaArray 8 $ \ po -> do
    _ <- cRotmg sd1 sd2 sx1 sy1 po nullPtr
    [flag, a, b, c, e11, e12, e21, e22] <- peekArray 8 po
    return $ case (round flag :: Int) of
        -2 -> FLAGNEG2
        -1 -> FLAGNEG1 { d1 = a, d2 = b, x1 = c, h11 = e11, h12 = e12, h21 = e21, h22 = e22 }
        0  -> FLAG0    { d1 = a, d2 = b, x1 = c, h12 = e12, h21 = e21 }
        _  -> FLAG1    { d1 = a, d2 = b, x1 = c, h11 = e11, h22 = e22 }

-- ---------------------------------------------------------------------------
-- the reference's deprecated monomorphic aliases (Single.hs:83-87, 112-116, 128-132,
-- 146-150, 164-168, 200-204, 221-225, 297-300, 354-357, 420-424, 440-444, 483-487)
-- ---------------------------------------------------------------------------
sdot :: Int -> DVector Float -> Int -> DVector Float -> Int -> IO Float
sdot = dot
ddot :: Int -> DVector Double -> Int -> DVector Double -> Int -> IO Double
ddot = dot
sasum, snrm2 :: Int -> DVector Float -> Int -> IO Float
sasum = asum
snrm2 = nrm2
dasum, dnrm2 :: Int -> DVector Double -> Int -> IO Double
dasum = asum
dnrm2 = nrm2
isamax :: Int -> DVector Float -> Int -> IO Int
isamax = iamax
idamax :: Int -> DVector Double -> Int -> IO Int
idamax = iamax
sscal :: Int -> Float -> DVector Float -> Int -> IO (DVector Float)
sscal = scal
dscal :: Int -> Double -> DVector Double -> Int -> IO (DVector Double)
dscal = scal
scopy :: Int -> DVector Float -> Int -> DVector Float -> Int -> IO (DVector Float)
scopy = copy
dcopy :: Int -> DVector Double -> Int -> DVector Double -> Int -> IO (DVector Double)
dcopy = copy
sswap :: Int -> DVector Float -> Int -> DVector Float -> Int -> IO (DVector Float, DVector Float)
sswap = swap
dswap :: Int -> DVector Double -> Int -> DVector Double -> Int -> IO (DVector Double, DVector Double)
dswap = swap
saxpy :: Int -> Float -> DVector Float -> Int -> DVector Float -> Int -> IO (DVector Float)
saxpy = axpy
daxpy :: Int -> Double -> DVector Double -> Int -> DVector Double -> Int -> IO (DVector Double)
daxpy = axpy
srot :: Int -> DVector Float -> Int -> DVector Float -> Int -> Float -> Float -> IO (DVector Float, DVector Float)
srot = rot
drot :: Int -> DVector Double -> Int -> DVector Double -> Int -> Double -> Double -> IO (DVector Double, DVector Double)
drot = rot
srotm :: ModGivensRot Float -> Int -> DVector Float -> Int -> DVector Float -> Int -> IO (DVector Float, DVector Float)
srotm = rotm
drotm :: ModGivensRot Double -> Int -> DVector Double -> Int -> DVector Double -> Int -> IO (DVector Double, DVector Double)
drotm = rotm
srotg :: Float -> Float -> IO (GivensRot Float)
srotg = rotg
drotg :: Double -> Double -> IO (GivensRot Double)
drotg = rotg
srotmg :: Float -> Float -> Float -> Float -> IO (ModGivensRot Float)
srotmg = rotmg
drotmg :: Double -> Double -> Double -> Double -> IO (ModGivensRot Double)
drotmg = rotmg
