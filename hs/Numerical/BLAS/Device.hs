{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE BangPatterns #-}
{- | Numerical.BLAS.Device -- the single-precision Level-1 routines of
"Numerical.BLAS.Single" on an NVIDIA B200, behind the C ABI of @include/lbk.h@
(@lambda_blas_b200/liblbk.so@).

The call surface is the reference's (src/Numerical/BLAS/Single.hs:84,113,129,147,
165,201,222,421,441,484): same names, same argument order, same results -- with
@Vector Float@ replaced by the device-resident 'DVector' and the result in 'IO',
because the work is enqueued on a CUDA stream.  Like the reference the vector
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
    , DVector, dvLength, toDevice, fromDevice, cloneD
      -- * Products and norms
    , sdot, sdsdot, sasum, snrm2, isamax
      -- * Pure vector results (new device vectors)
    , sscal, scopy, sswap, saxpy, srot, srotm
      -- * In-place forms
    , saxpy_, sscal_, scopy_, sswap_, srot_, srotm_
      -- * Host scalars
    , srotg, srotmg
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
import Foreign.Storable (peek)
import Numerical.BLAS.Types (GivensRot(..), ModGivensRot(..))

-- ---------------------------------------------------------------------------
-- opaque handles
-- ---------------------------------------------------------------------------
data LbkCtx
data LbkVec

-- | One GPU, one stream, reduction scratch (and, after @lbk_comm_init@, a communicator).
newtype Context = Context (Ptr LbkCtx)

-- | A device-resident @Vector Float@: a finalised handle plus the context that owns it.
data DVector = DVector { dvCtx :: !Context, dvPtr :: !(ForeignPtr LbkVec), dvLength :: !Int }

data LbkError = LbkError Int String deriving Show
instance Exception LbkError

-- ---------------------------------------------------------------------------
-- the C ABI (include/lbk.h).  `unsafe` only for calls that merely enqueue or are
-- O(1) on the host; anything that blocks on the GPU is a safe call
-- (test/Foreign.hs binds both flavours; test/Bench.hs:122-126 times both).
-- ---------------------------------------------------------------------------
foreign import ccall        "lbk_init"        c_init        :: CInt -> Ptr (Ptr LbkCtx) -> IO CInt
foreign import ccall        "lbk_shutdown"    c_shutdown    :: Ptr LbkCtx -> IO CInt
foreign import ccall        "lbk_sync"        c_sync        :: Ptr LbkCtx -> IO CInt
foreign import ccall unsafe "lbk_last_error"  c_last_error  :: Ptr LbkCtx -> IO CString
foreign import ccall        "lbk_vec_alloc"   c_vec_alloc   :: Ptr LbkCtx -> Int -> Ptr (Ptr LbkVec) -> IO CInt
foreign import ccall        "&lbk_vec_free"   p_vec_free    :: FunPtr (Ptr LbkVec -> IO ())
foreign import ccall        "lbk_vec_upload"  c_vec_upload  :: Ptr LbkVec -> Ptr Float -> Int -> IO CInt
foreign import ccall        "lbk_vec_download" c_vec_download :: Ptr LbkVec -> Ptr Float -> Int -> IO CInt
foreign import ccall        "lbk_vec_clone"   c_vec_clone   :: Ptr LbkVec -> Ptr (Ptr LbkVec) -> IO CInt

foreign import ccall "lbk_sdot"   c_sdot   :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_sdsdot" c_sdsdot :: Ptr LbkCtx -> Int -> Float -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_sasum"  c_sasum  :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_snrm2"  c_snrm2  :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr Float -> IO CInt
foreign import ccall "lbk_isamax" c_isamax :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr Int -> IO CInt

foreign import ccall unsafe "lbk_saxpy" c_saxpy :: Ptr LbkCtx -> Int -> Float -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> IO CInt
foreign import ccall unsafe "lbk_sscal" c_sscal :: Ptr LbkCtx -> Int -> Float -> Ptr LbkVec -> Int -> IO CInt
foreign import ccall unsafe "lbk_scopy" c_scopy :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> IO CInt
foreign import ccall unsafe "lbk_sswap" c_sswap :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> IO CInt
foreign import ccall unsafe "lbk_srot"  c_srot  :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Float -> Float -> IO CInt
foreign import ccall unsafe "lbk_srotm" c_srotm :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr Float -> IO CInt

foreign import ccall unsafe "lbk_saxpy_oop" c_saxpy_oop :: Ptr LbkCtx -> Int -> Float -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> IO CInt
foreign import ccall unsafe "lbk_sscal_oop" c_sscal_oop :: Ptr LbkCtx -> Int -> Float -> Ptr LbkVec -> Int -> Ptr LbkVec -> IO CInt
foreign import ccall unsafe "lbk_scopy_oop" c_scopy_oop :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> IO CInt
foreign import ccall unsafe "lbk_sswap_oop" c_sswap_oop :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Ptr LbkVec -> IO CInt
foreign import ccall unsafe "lbk_srot_oop"  c_srot_oop  :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Float -> Float -> Ptr LbkVec -> Ptr LbkVec -> IO CInt
foreign import ccall unsafe "lbk_srotm_oop" c_srotm_oop :: Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr LbkVec -> Int -> Ptr Float -> Ptr LbkVec -> Ptr LbkVec -> IO CInt

foreign import ccall unsafe "lbk_srotg"  c_srotg  :: Float -> Float -> Ptr Float -> IO CInt
foreign import ccall unsafe "lbk_srotmg" c_srotmg :: Float -> Float -> Float -> Float -> Ptr Float -> Ptr Float -> IO CInt

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

adopt :: Context -> Int -> Ptr LbkVec -> IO DVector
adopt ctx n p = do fp <- newForeignPtr p_vec_free p
                   return (DVector ctx fp n)

allocD :: Context -> Int -> IO DVector
allocD ctx@(Context c) n = alloca $ \ pp -> do
    c_vec_alloc c n pp >>= check ctx
    peek pp >>= adopt ctx n

-- | Explicit host -> device transfer (the north star's "explicit host transfer").
toDevice :: Context -> V.Vector Float -> IO DVector
toDevice ctx v = do
    d <- allocD ctx (V.length v)
    V.unsafeWith v $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        c_vec_upload dp hp (V.length v) >>= check ctx
    return d

-- | Explicit device -> host transfer.
fromDevice :: DVector -> IO (V.Vector Float)
fromDevice d = do
    mv <- VM.new (dvLength d)
    VM.unsafeWith mv $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        c_vec_download dp hp (dvLength d) >>= check (dvCtx d)
    V.unsafeFreeze mv

cloneD :: DVector -> IO DVector
cloneD d = alloca $ \ pp -> withForeignPtr (dvPtr d) $ \ dp -> do
    c_vec_clone dp pp >>= check (dvCtx d)
    peek pp >>= adopt (dvCtx d) (dvLength d)

with2 :: DVector -> DVector -> (Ptr LbkCtx -> Ptr LbkVec -> Ptr LbkVec -> IO a) -> IO a
with2 x y f = let Context c = dvCtx x in
    withForeignPtr (dvPtr x) $ \ px -> withForeignPtr (dvPtr y) $ \ py -> f c px py

-- ---------------------------------------------------------------------------
-- reductions (Single.hs:84,165,201,222,231)
-- ---------------------------------------------------------------------------
sdot :: Int -> DVector -> Int -> DVector -> Int -> IO Float
sdot n x incx y incy = alloca $ \ po -> with2 x y $ \ c px py -> do
    c_sdot c n px incx py incy po >>= check (dvCtx x)
    peek po

sdsdot :: Int -> Float -> DVector -> Int -> DVector -> Int -> IO Float
sdsdot n sb x incx y incy = alloca $ \ po -> with2 x y $ \ c px py -> do
    c_sdsdot c n sb px incx py incy po >>= check (dvCtx x)
    peek po

ivi :: (Ptr LbkCtx -> Int -> Ptr LbkVec -> Int -> Ptr Float -> IO CInt) -> Int -> DVector -> Int -> IO Float
ivi f n x incx = let Context c = dvCtx x in
    alloca $ \ po -> withForeignPtr (dvPtr x) $ \ px -> do
        f c n px incx po >>= check (dvCtx x)
        peek po

sasum, snrm2 :: Int -> DVector -> Int -> IO Float
sasum = ivi c_sasum
snrm2 = ivi c_snrm2

-- | 0-based logical index, -1 when n<1 or incx<1 (Single.hs:217-220).
isamax :: Int -> DVector -> Int -> IO Int
isamax n x incx = let Context c = dvCtx x in
    alloca $ \ po -> withForeignPtr (dvPtr x) $ \ px -> do
        c_isamax c n px incx po >>= check (dvCtx x)
        peek po

-- ---------------------------------------------------------------------------
-- pure vector results: new device vectors, arguments untouched (Util.hs:134-142)
-- ---------------------------------------------------------------------------
sscal :: Int -> Float -> DVector -> Int -> IO DVector
sscal n a x incx = do
    o <- allocD (dvCtx x) (dvLength x)
    with2 x o $ \ c px po -> c_sscal_oop c n a px incx po >>= check (dvCtx x)
    return o

scopy :: Int -> DVector -> Int -> DVector -> Int -> IO DVector
scopy n x incx y incy = do
    o <- allocD (dvCtx y) (dvLength y)
    with2 x y $ \ c px py -> withForeignPtr (dvPtr o) $ \ po ->
        c_scopy_oop c n px incx py incy po >>= check (dvCtx x)
    return o

saxpy :: Int -> Float -> DVector -> Int -> DVector -> Int -> IO DVector
saxpy n a x incx y incy = do
    o <- allocD (dvCtx y) (dvLength y)
    with2 x y $ \ c px py -> withForeignPtr (dvPtr o) $ \ po ->
        c_saxpy_oop c n a px incx py incy po >>= check (dvCtx x)
    return o

pair :: DVector -> DVector -> (Ptr LbkCtx -> Ptr LbkVec -> Ptr LbkVec -> Ptr LbkVec -> Ptr LbkVec -> IO CInt) -> IO (DVector, DVector)
pair x y f = do
    xo <- allocD (dvCtx x) (dvLength x)
    yo <- allocD (dvCtx y) (dvLength y)
    with2 x y $ \ c px py -> with2 xo yo $ \ _ pxo pyo -> f c px py pxo pyo >>= check (dvCtx x)
    return (xo, yo)

sswap :: Int -> DVector -> Int -> DVector -> Int -> IO (DVector, DVector)
sswap n x incx y incy = pair x y $ \ c px py pxo pyo -> c_sswap_oop c n px incx py incy pxo pyo

srot :: Int -> DVector -> Int -> DVector -> Int -> Float -> Float -> IO (DVector, DVector)
srot n x incx y incy cs sn = pair x y $ \ c px py pxo pyo -> c_srot_oop c n px incx py incy cs sn pxo pyo

-- | BLAS sparam [flag,h11,h21,h12,h22], filled exactly as test/Foreign.hs:276-280 does.
sparam :: ModGivensRot Float -> [Float]
sparam FLAGNEG2 = [-2, 1, 0, 0, 1]
sparam m@FLAGNEG1{} = [-1, h11 m, h21 m, h12 m, h22 m]
sparam m@FLAG0{}    = [0, 0, h21 m, h12 m, 0]
sparam m@FLAG1{}    = [1, h11 m, 0, 0, h22 m]

-- | FLAGNEG2 hands back the arguments themselves, as the reference does (Single.hs:473).
srotm :: ModGivensRot Float -> Int -> DVector -> Int -> DVector -> Int -> IO (DVector, DVector)
srotm FLAGNEG2 _ x _ y _ = return (x, y)
srotm flags n x incx y incy = withArray (sparam flags) $ \ pp ->
    pair x y $ \ c px py pxo pyo -> c_srotm_oop c n px incx py incy pp pxo pyo

-- ---------------------------------------------------------------------------
-- in-place forms (enqueue only)
-- ---------------------------------------------------------------------------
saxpy_ :: Int -> Float -> DVector -> Int -> DVector -> Int -> IO ()
saxpy_ n a x incx y incy = with2 x y $ \ c px py -> c_saxpy c n a px incx py incy >>= check (dvCtx x)

sscal_ :: Int -> Float -> DVector -> Int -> IO ()
sscal_ n a x incx = let Context c = dvCtx x in
    withForeignPtr (dvPtr x) $ \ px -> c_sscal c n a px incx >>= check (dvCtx x)

scopy_, sswap_ :: Int -> DVector -> Int -> DVector -> Int -> IO ()
scopy_ n x incx y incy = with2 x y $ \ c px py -> c_scopy c n px incx py incy >>= check (dvCtx x)
sswap_ n x incx y incy = with2 x y $ \ c px py -> c_sswap c n px incx py incy >>= check (dvCtx x)

srot_ :: Int -> DVector -> Int -> DVector -> Int -> Float -> Float -> IO ()
srot_ n x incx y incy cs sn = with2 x y $ \ c px py -> c_srot c n px incx py incy cs sn >>= check (dvCtx x)

srotm_ :: ModGivensRot Float -> Int -> DVector -> Int -> DVector -> Int -> IO ()
srotm_ flags n x incx y incy = withArray (sparam flags) $ \ pp ->
    with2 x y $ \ c px py -> c_srotm c n px incx py incy pp >>= check (dvCtx x)

-- ---------------------------------------------------------------------------
-- host scalars (Single.hs:274-300, 318-357); pure, like the reference's
-- ---------------------------------------------------------------------------
srotg :: Float -> Float -> IO (GivensRot Float)
srotg a b = allocaArray 4 $ \ po -> do
    _ <- c_srotg a b po
    [r, z, c, s] <- peekArray 4 po
    return (GIVENSROT (r, z, c, s))

srotmg :: Float -> Float -> Float -> Float -> IO (ModGivensRot Float)
srotmg sd1 sd2 sx1 sy1 = allocaArray 8 $ \ po -> do
    _ <- c_srotmg sd1 sd2 sx1 sy1 po nullPtr
    [flag, a, b, c, e11, e12, e21, e22] <- peekArray 8 po
    return $ case (round flag :: Int) of
        -2 -> FLAGNEG2
        -1 -> FLAGNEG1 { d1 = a, d2 = b, x1 = c, h11 = e11, h12 = e12, h21 = e21, h22 = e22 }
        0  -> FLAG0    { d1 = a, d2 = b, x1 = c, h12 = e12, h21 = e21 }
        _  -> FLAG1    { d1 = a, d2 = b, x1 = c, h11 = e11, h22 = e22 }
