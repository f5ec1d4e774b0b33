{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE BangPatterns #-}
{- | Numerical.BLAS.Device -- the Level-1 routines of "Numerical.BLAS.Single" on an
NVIDIA B200, behind the C ABI of @include/lbk.h@ (@lambda_blas_b200/liblbk.so@).

The call surface is the reference's (src/Numerical/BLAS/Single.hs:41-61): the
polymorphic names (@dot@, @axpy@, ...) with their deprecated aliases for Float
(@sdot@, ..., Single.hs:84,113,129,147,165,201,222,421,441,484) and Double (@ddot@,
..., Single.hs:86,115,131,149,167,203,224,423,443,486): same argument order, same
results -- with @Vector a@ replaced by the device-resident 'DVector' @a@ (defined in
"Numerical.BLAS.Device.Types", re-exported by the patched "Numerical.BLAS.Types") and the
result in 'IO', because the work is enqueued on a CUDA stream.  The O(1) host
scalars 'rotg' / 'rotmg' stay pure, exactly as in the reference (Single.hs:274,318).  Where the reference
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
    , Elt, DVector, dvLength, dvLayout, Layout(..), toDevice, fromDevice, cloneD
      -- * One host thread, many GPUs: sharded vectors (lbk_init_multi)
    , withContextMulti, newContextMulti, toDeviceSharded, allocSharded, setMultiThreads, setFusedExchange
    , setExchangeTimeout
      -- * Recorded chains of calls (CUDA graphs)
    , Graph, capture, launchGraph
      -- * Asynchronous and pinned transfers
    , withRegistered, toDeviceAsync, fromDeviceAsync
      -- * Enqueue-only reductions (the result lands in a device vector)
    , dotAsync, asumAsync, nrm2Async, iamaxAsync
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

import Control.Exception (bracket, bracket_, onException, throwIO)
import Control.Monad (when)
import Data.Word (Word8)
import qualified Data.Vector.Storable as V
import qualified Data.Vector.Storable.Mutable as VM
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types
import Foreign.ForeignPtr
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, withArray)
import Foreign.Ptr
import Foreign.Storable (Storable, peek, sizeOf)
import System.IO.Unsafe (unsafePerformIO)
import Numerical.BLAS.Device.Types
import Numerical.BLAS.Types (GivensRot(..), ModGivensRot(..))

-- ---------------------------------------------------------------------------
-- the C ABI (include/lbk.h).  `unsafe` only for calls that merely enqueue or are
-- O(1) on the host; anything that blocks on the GPU is a safe call
-- (test/Foreign.hs binds both flavours; test/Bench.hs:122-126 times both).
-- ---------------------------------------------------------------------------
type V a = Ptr (LbkVec a)

foreign import ccall        "lbk_init"        c_init        :: CInt -> Ptr (Ptr LbkCtx) -> IO CInt
foreign import ccall        "lbk_init_multi"  c_init_multi  :: CInt -> Ptr CInt -> Ptr (Ptr LbkCtx) -> IO CInt
foreign import ccall unsafe "lbk_set_multi_threads"    c_set_multi_threads    :: Ptr LbkCtx -> CInt -> IO CInt
foreign import ccall unsafe "lbk_set_fused_exchange"   c_set_fused_exchange   :: Ptr LbkCtx -> CInt -> IO CInt
foreign import ccall unsafe "lbk_set_exchange_timeout" c_set_exchange_timeout :: Ptr LbkCtx -> Int -> IO CInt
foreign import ccall        "lbk_host_register"   c_host_register   :: Ptr Word8 -> Int -> IO CInt
foreign import ccall        "lbk_host_unregister" c_host_unregister :: Ptr Word8 -> IO CInt
foreign import ccall        "lbk_shutdown"    c_shutdown    :: Ptr LbkCtx -> IO CInt
foreign import ccall        "lbk_sync"        c_sync        :: Ptr LbkCtx -> IO CInt
foreign import ccall unsafe "lbk_last_error"  c_last_error  :: Ptr LbkCtx -> IO CString
foreign import ccall        "&lbk_vec_free"   p_vec_free    :: FunPtr (Ptr () -> IO ())
foreign import ccall        "lbk_capture_begin" c_capture_begin :: Ptr LbkCtx -> IO CInt
foreign import ccall        "lbk_capture_end"   c_capture_end   :: Ptr LbkCtx -> Ptr (Ptr LbkGraph) -> IO CInt
foreign import ccall unsafe "lbk_graph_launch"  c_graph_launch  :: Ptr LbkGraph -> IO CInt
foreign import ccall        "&lbk_graph_destroy" p_graph_destroy :: FunPtr (Ptr LbkGraph -> IO ())

-- Float instance (s / is prefix)
foreign import ccall        "lbk_vec_alloc"    s_vec_alloc    :: Ptr LbkCtx -> Int -> Ptr (V Float) -> IO CInt
foreign import ccall        "lbk_vec_upload"   s_vec_upload   :: V Float -> Ptr Float -> Int -> IO CInt
foreign import ccall        "lbk_vec_download" s_vec_download :: V Float -> Ptr Float -> Int -> IO CInt
foreign import ccall        "lbk_vec_clone"    s_vec_clone    :: V Float -> Ptr (V Float) -> IO CInt
foreign import ccall        "lbk_vec_alloc_sharded" s_vec_alloc_sharded :: Ptr LbkCtx -> Int -> Ptr (V Float) -> IO CInt
foreign import ccall        "lbk_vec_alloc_like"    s_vec_alloc_like    :: V Float -> Ptr (V Float) -> IO CInt
foreign import ccall unsafe "lbk_vec_upload_async"   s_vec_upload_async   :: V Float -> Ptr Float -> Int -> IO CInt
foreign import ccall unsafe "lbk_vec_download_async" s_vec_download_async :: V Float -> Ptr Float -> Int -> IO CInt
foreign import ccall unsafe "lbk_sdot_async"   s_dot_async   :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> Int -> V Float -> IO CInt
foreign import ccall unsafe "lbk_sasum_async"  s_asum_async  :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> IO CInt
foreign import ccall unsafe "lbk_snrm2_async"  s_nrm2_async  :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> IO CInt
foreign import ccall unsafe "lbk_isamax_async" s_iamax_async :: Ptr LbkCtx -> Int -> V Float -> Int -> V Float -> IO CInt
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
foreign import ccall        "lbk_vec_alloc_sharded_f64" d_vec_alloc_sharded :: Ptr LbkCtx -> Int -> Ptr (V Double) -> IO CInt
foreign import ccall        "lbk_vec_alloc_like"    d_vec_alloc_like    :: V Double -> Ptr (V Double) -> IO CInt
foreign import ccall unsafe "lbk_vec_upload_async"   d_vec_upload_async   :: V Double -> Ptr Double -> Int -> IO CInt
foreign import ccall unsafe "lbk_vec_download_async" d_vec_download_async :: V Double -> Ptr Double -> Int -> IO CInt
foreign import ccall unsafe "lbk_ddot_async"   d_dot_async   :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> Int -> V Double -> IO CInt
foreign import ccall unsafe "lbk_dasum_async"  d_asum_async  :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> IO CInt
foreign import ccall unsafe "lbk_dnrm2_async"  d_nrm2_async  :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> IO CInt
foreign import ccall unsafe "lbk_idamax_async" d_iamax_async :: Ptr LbkCtx -> Int -> V Double -> Int -> V Double -> IO CInt
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
    cVecAllocSharded  :: Ptr LbkCtx -> Int -> Ptr (V a) -> IO CInt
    cVecAllocLike     :: V a -> Ptr (V a) -> IO CInt
    cVecUploadAsync   :: V a -> Ptr a -> Int -> IO CInt
    cVecDownloadAsync :: V a -> Ptr a -> Int -> IO CInt
    cDotAsync   :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> Int -> V a -> IO CInt
    cAsumAsync  :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> IO CInt
    cNrm2Async  :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> IO CInt
    cIamaxAsync :: Ptr LbkCtx -> Int -> V a -> Int -> V a -> IO CInt
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
    cVecAllocSharded = s_vec_alloc_sharded; cVecAllocLike = s_vec_alloc_like
    cVecUploadAsync = s_vec_upload_async; cVecDownloadAsync = s_vec_download_async
    cDotAsync = s_dot_async; cAsumAsync = s_asum_async; cNrm2Async = s_nrm2_async; cIamaxAsync = s_iamax_async
    cDot = s_dot; cAsum = s_asum; cNrm2 = s_nrm2; cIamax = s_iamax
    cAxpy = s_axpy; cScal = s_scal; cCopy = s_copy; cSwap = s_swap; cRot = s_rot; cRotm = s_rotm
    cAxpyOop = s_axpy_oop; cScalOop = s_scal_oop; cCopyOop = s_copy_oop; cSwapOop = s_swap_oop
    cRotOop = s_rot_oop; cRotmOop = s_rotm_oop; cRotg = s_rotg; cRotmg = s_rotmg

instance Elt Double where
    cVecAlloc = d_vec_alloc; cVecUpload = d_vec_upload; cVecDownload = d_vec_download; cVecClone = d_vec_clone
    cVecAllocSharded = d_vec_alloc_sharded; cVecAllocLike = d_vec_alloc_like
    cVecUploadAsync = d_vec_upload_async; cVecDownloadAsync = d_vec_download_async
    cDotAsync = d_dot_async; cAsumAsync = d_asum_async; cNrm2Async = d_nrm2_async; cIamaxAsync = d_iamax_async
    cDot = d_dot; cAsum = d_asum; cNrm2 = d_nrm2; cIamax = d_iamax
    cAxpy = d_axpy; cScal = d_scal; cCopy = d_copy; cSwap = d_swap; cRot = d_rot; cRotm = d_rotm
    cAxpyOop = d_axpy_oop; cScalOop = d_scal_oop; cCopyOop = d_copy_oop; cSwapOop = d_swap_oop
    cRotOop = d_rot_oop; cRotmOop = d_rotm_oop; cRotg = d_rotg; cRotmg = d_rotmg

-- ---------------------------------------------------------------------------
-- helpers
-- ---------------------------------------------------------------------------
check :: Context -> CInt -> IO ()
check ctx st = when (st /= 0) $ do
    msg <- c_last_error (ctxPtr ctx) >>= peekCString
    throwIO (LbkError (fromIntegral st) msg)

checkNoCtx :: CInt -> IO ()
checkNoCtx st = when (st /= 0) $ c_last_error nullPtr >>= peekCString >>= throwIO . LbkError (fromIntegral st)

newContext :: Int -> IO Context
newContext dev = alloca $ \ pp -> do
    c_init (fromIntegral dev) pp >>= checkNoCtx
    p <- peek pp
    return (Context p 1)

-- | ONE host thread, many GPUs (@lbk_init_multi@): rank r is @devs !! r@.  'toDeviceSharded' then makes vectors
-- that are block-sharded over the ranks, and every routine of this module fans out over the devices; a
-- reduction returns the scalar all ranks agree on.  'toDevice' still makes whole vectors (on the first device).
newContextMulti :: [Int] -> IO Context
newContextMulti devs = alloca $ \ pp -> withArray (map fromIntegral devs) $ \ pd -> do
    c_init_multi (fromIntegral (length devs)) pd pp >>= checkNoCtx
    p <- peek pp
    return (Context p (length devs))

-- | @lbk_shutdown@.  Device vectors that are still reachable keep the native context alive until their
-- finalisers have run (the C library counts handles), so closing first and collecting later is safe.
closeContext :: Context -> IO ()
closeContext ctx = () <$ c_shutdown (ctxPtr ctx)

withContext :: Int -> (Context -> IO a) -> IO a
withContext dev = bracket (newContext dev) closeContext

withContextMulti :: [Int] -> (Context -> IO a) -> IO a
withContextMulti devs = bracket (newContextMulti devs) closeContext

sync :: Context -> IO ()
sync ctx = c_sync (ctxPtr ctx) >>= check ctx

-- | One launcher thread per device ('True', the default on distinct devices) or rank after rank from the calling thread.
setMultiThreads :: Context -> Bool -> IO ()
setMultiThreads ctx on = c_set_multi_threads (ctxPtr ctx) (if on then 1 else 0) >>= check ctx

-- | In-kernel record exchange over peer memory ('True') or the stream-ordered one (records, events, a finish kernel).
setFusedExchange :: Context -> Bool -> IO ()
setFusedExchange ctx on = c_set_fused_exchange (ctxPtr ctx) (if on then 1 else 0) >>= check ctx

-- | Milliseconds a sharded reduction waits for its peers' records before it fails with @LBK_ERR_NCCL@.
setExchangeTimeout :: Context -> Int -> IO ()
setExchangeTimeout ctx ms = c_set_exchange_timeout (ctxPtr ctx) ms >>= check ctx

-- | A recorded chain of calls (@lbk_graph@).  Opaque; the finaliser is @lbk_graph_destroy@.
data LbkGraph
data Graph = Graph !Context !(ForeignPtr LbkGraph)

-- | @capture ctx act@ RECORDS what @act@ enqueues on @ctx@ instead of running it (@lbk_capture_begin@ /
-- @lbk_capture_end@): the in-place forms ('axpy_', 'scal_', ...), the enqueue-only reductions ('dotAsync', ...) and the
-- asynchronous transfers.  Whatever cannot be recorded -- the pure forms (they allocate), scalars returned to the
-- host, 'sync' -- throws 'LbkError' with @LBK_ERR_UNSUPPORTED@ and leaves the recording intact.  'launchGraph' replays
-- the chain with one host call: the regime of the reference's own benchmark (test/Bench.hs:113, n <= 30 000) is
-- bound by launch latency, not by bandwidth.
capture :: Context -> IO () -> IO Graph
capture ctx act = alloca $ \ pg -> do
    c_capture_begin (ctxPtr ctx) >>= check ctx
    act `onException` c_capture_end (ctxPtr ctx) pg
    c_capture_end (ctxPtr ctx) pg >>= check ctx
    g <- peek pg
    fp <- newForeignPtr p_graph_destroy g
    return (Graph ctx fp)

-- | One replay of a recorded chain, enqueue-only (@lbk_graph_launch@).
launchGraph :: Graph -> IO ()
launchGraph (Graph ctx fp) = withForeignPtr fp $ \ g -> c_graph_launch g >>= check ctx

adopt :: Context -> Int -> Layout -> V a -> IO (DVector a)
adopt ctx n lay p = do fp <- newForeignPtr (castFunPtr p_vec_free) p
                       return (DVector ctx fp n lay)

allocD :: Elt a => Context -> Int -> IO (DVector a)
allocD ctx n = alloca $ \ pp -> do
    cVecAlloc (ctxPtr ctx) n pp >>= check ctx
    peek pp >>= adopt ctx n Whole

-- | An uninitialised vector of logical length n, block-sharded over the ranks of a multi-device context.
allocSharded :: Elt a => Context -> Int -> IO (DVector a)
allocSharded ctx n = alloca $ \ pp -> do
    cVecAllocSharded (ctxPtr ctx) n pp >>= check ctx
    peek pp >>= adopt ctx n (if ctxRanks ctx > 1 then Sharded else Whole)

-- | An uninitialised vector with the length, element type and layout of its argument: what a pure result is made of.
allocLike :: Elt a => DVector a -> IO (DVector a)
allocLike d = alloca $ \ pp -> withForeignPtr (dvPtr d) $ \ dp -> do
    cVecAllocLike dp pp >>= check (dvCtx d)
    peek pp >>= adopt (dvCtx d) (dvLength d) (dvLayout d)

-- | Explicit host -> device transfer (the north star's "explicit host transfer").
toDevice :: Elt a => Context -> V.Vector a -> IO (DVector a)
toDevice ctx v = do
    d <- allocD ctx (V.length v)
    V.unsafeWith v $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        cVecUpload dp hp (V.length v) >>= check ctx
    return d

-- | Explicit host -> devices transfer of ONE host vector into block shards: every rank uploads its own block on its
-- own stream (one PCIe link each).
toDeviceSharded :: Elt a => Context -> V.Vector a -> IO (DVector a)
toDeviceSharded ctx v = do
    d <- allocSharded ctx (V.length v)
    V.unsafeWith v $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        cVecUpload dp hp (V.length v) >>= check ctx
    return d

-- | Explicit device -> host transfer (of the whole logical vector: shards are gathered into one host buffer).
fromDevice :: Elt a => DVector a -> IO (V.Vector a)
fromDevice d = do
    mv <- VM.new (dvLength d)
    VM.unsafeWith mv $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
        cVecDownload dp hp (dvLength d) >>= check (dvCtx d)
    V.unsafeFreeze mv

-- | Page-locks the buffer of a host vector for the duration of the action (@lbk_host_register@): uploads from it
-- ('toDevice', 'toDeviceAsync') then run at the rate of a pinned copy instead of through the driver's staging buffer.
withRegistered :: Storable a => V.Vector a -> IO b -> IO b
withRegistered v act
    | V.null v  = act
    | otherwise = V.unsafeWith v $ \ hp ->
        let p = castPtr hp :: Ptr Word8
            nbytes = V.length v * sizeOf (V.head v)
        in bracket_ (c_host_register p nbytes >>= checkNoCtx) (c_host_unregister p) act

-- | Enqueue-only upload into an existing device vector; the host vector must stay alive (and should be registered)
-- until 'sync'.
toDeviceAsync :: Elt a => DVector a -> V.Vector a -> IO ()
toDeviceAsync d v = V.unsafeWith v $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
    cVecUploadAsync dp hp (min (V.length v) (dvLength d)) >>= check (dvCtx d)

-- | Enqueue-only download into a mutable host vector; valid after 'sync'.
fromDeviceAsync :: Elt a => DVector a -> VM.IOVector a -> IO ()
fromDeviceAsync d mv = VM.unsafeWith mv $ \ hp -> withForeignPtr (dvPtr d) $ \ dp ->
    cVecDownloadAsync dp hp (min (VM.length mv) (dvLength d)) >>= check (dvCtx d)

cloneD :: Elt a => DVector a -> IO (DVector a)
cloneD d = alloca $ \ pp -> withForeignPtr (dvPtr d) $ \ dp -> do
    cVecClone dp pp >>= check (dvCtx d)
    peek pp >>= adopt (dvCtx d) (dvLength d) (dvLayout d)

with2 :: DVector a -> DVector a -> (Ptr LbkCtx -> V a -> V a -> IO b) -> IO b
with2 x y f = withForeignPtr (dvPtr x) $ \ px -> withForeignPtr (dvPtr y) $ \ py -> f (ctxPtr (dvCtx x)) px py

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
ivi f n x incx = let c = ctxPtr (dvCtx x) in
    alloca $ \ po -> withForeignPtr (dvPtr x) $ \ px -> do
        f c n px incx po >>= check (dvCtx x)
        peek po

asum, nrm2 :: Elt a => Int -> DVector a -> Int -> IO a
asum = ivi cAsum
nrm2 = ivi cNrm2

-- | 0-based logical index, -1 when n<1 or incx<1 (Single.hs:217-220).
iamax :: Elt a => Int -> DVector a -> Int -> IO Int
iamax n x incx = let c = ctxPtr (dvCtx x) in
    alloca $ \ po -> withForeignPtr (dvPtr x) $ \ px -> do
        cIamax c n px incx po >>= check (dvCtx x)
        peek po

-- | Enqueue-only forms: the scalar lands in element 0 of @out@ (for 'iamaxAsync': the 64-bit index in its first
-- 8 bytes), a 'Whole' device vector of the routine's element type; nothing is synchronised.
dotAsync :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> DVector a -> IO ()
dotAsync n x incx y incy out = with2 x y $ \ c px py -> withForeignPtr (dvPtr out) $ \ po ->
    cDotAsync c n px incx py incy po >>= check (dvCtx x)

iviAsync :: (Ptr LbkCtx -> Int -> V a -> Int -> V a -> IO CInt) -> Int -> DVector a -> Int -> DVector a -> IO ()
iviAsync f n x incx out = with2 x out $ \ c px po -> f c n px incx po >>= check (dvCtx x)

asumAsync, nrm2Async, iamaxAsync :: Elt a => Int -> DVector a -> Int -> DVector a -> IO ()
asumAsync = iviAsync cAsumAsync
nrm2Async = iviAsync cNrm2Async
iamaxAsync = iviAsync cIamaxAsync

-- ---------------------------------------------------------------------------
-- pure vector results: new device vectors, arguments untouched (Util.hs:134-142)
-- ---------------------------------------------------------------------------
scal :: Elt a => Int -> a -> DVector a -> Int -> IO (DVector a)
scal n a x incx = do
    o <- allocLike x
    with2 x o $ \ c px po -> cScalOop c n a px incx po >>= check (dvCtx x)
    return o

copy :: Elt a => Int -> DVector a -> Int -> DVector a -> Int -> IO (DVector a)
copy n x incx y incy = do
    o <- allocLike y
    with2 x y $ \ c px py -> withForeignPtr (dvPtr o) $ \ po ->
        cCopyOop c n px incx py incy po >>= check (dvCtx x)
    return o

axpy :: Elt a => Int -> a -> DVector a -> Int -> DVector a -> Int -> IO (DVector a)
axpy n a x incx y incy = do
    o <- allocLike y
    with2 x y $ \ c px py -> withForeignPtr (dvPtr o) $ \ po ->
        cAxpyOop c n a px incx py incy po >>= check (dvCtx x)
    return o

pair :: Elt a => DVector a -> DVector a -> (Ptr LbkCtx -> V a -> V a -> V a -> V a -> IO CInt) -> IO (DVector a, DVector a)
pair x y f = do
    xo <- allocLike x
    yo <- allocLike y
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
scal_ n a x incx = let c = ctxPtr (dvCtx x) in
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
rotg :: Elt a => a -> a -> GivensRot a
rotg a b = unsafePerformIO $ allocaArray 4 $ \ po -> do      -- a pure C function of its arguments (lbk_scalars.cpp)
    _ <- cRotg a b po
    [r, z, c, s] <- peekArray 4 po
    return (GIVENSROT (r, z, c, s))
{-# NOINLINE rotg #-}

rotmg :: Elt a => a -> a -> a -> a -> ModGivensRot a
rotmg sd1 sd2 sx1 sy1 = unsafePerformIO $ allocaArray 8 $ \ po -> do
    _ <- cRotmg sd1 sd2 sx1 sy1 po nullPtr
    [flag, a, b, c, e11, e12, e21, e22] <- peekArray 8 po
    return $ case (round flag :: Int) of
        -2 -> FLAGNEG2
        -1 -> FLAGNEG1 { d1 = a, d2 = b, x1 = c, h11 = e11, h12 = e12, h21 = e21, h22 = e22 }
        0  -> FLAG0    { d1 = a, d2 = b, x1 = c, h12 = e12, h21 = e21 }
        _  -> FLAG1    { d1 = a, d2 = b, x1 = c, h11 = e11, h22 = e22 }
{-# NOINLINE rotmg #-}

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
srotg :: Float -> Float -> GivensRot Float
srotg = rotg
drotg :: Double -> Double -> GivensRot Double
drotg = rotg
srotmg :: Float -> Float -> Float -> Float -> ModGivensRot Float
srotmg = rotmg
drotmg :: Double -> Double -> Double -> Double -> ModGivensRot Double
drotmg = rotmg
