{- | Numerical.BLAS.Device.Types -- the device-resident vector representation that the B200 path adds beside the
reference's ADTs of "Numerical.BLAS.Types" (src/Numerical/BLAS/Types.hs:3 exports @GivensRot(..)@ and
@ModGivensRot(..)@ only; vectors in the reference are host buffers, @Data.Vector.Storable.Vector a@,
Single.hs:67).

@hs/patches/Types.hs.patch@ makes "Numerical.BLAS.Types" re-export this module, so that a program importing the
reference's type module sees the host vector's counterpart there, as the north star puts it ("the vector type gains a
device-buffer representation with explicit host transfer"); @hs/patches/lambda-blas.cabal.patch@ lists the new modules
and links @liblbk@.

NOTE: no GHC in this image (SURVEY.md, Appendix C): delivered as source, not compiled here.
-}
module Numerical.BLAS.Device.Types
    ( LbkCtx, LbkVec, Context(..), DVector(..), Layout(..), LbkError(..)
    ) where

import Control.Exception (Exception)
import Foreign.ForeignPtr (ForeignPtr)
import Foreign.Ptr (Ptr)

-- | Opaque C handles of include/lbk.h.
data LbkCtx
data LbkVec a          -- phantom: the element type of the device buffer

-- | One GPU (@lbk_init@) or, from ONE host thread, several (@lbk_init_multi@): streams, reduction scratch, the
-- mailboxes the ranks exchange their reduction records through.  'ctxRanks' is 1 for a one-device context.
data Context = Context { ctxPtr :: !(Ptr LbkCtx), ctxRanks :: !Int }

-- | Where a device vector lives: whole on one device, or block-sharded over the ranks of a multi-device context
-- (rank r holds the r-th contiguous block; @lbk_shard_range@).
data Layout = Whole | Sharded deriving (Eq, Show)

-- | A device-resident @Vector a@: a finalised handle, the context that owns it, its logical length and layout.
-- The finaliser is @lbk_vec_free@; the C library keeps a context alive until its last vector is freed, so a
-- finaliser that the garbage collector runs after 'Numerical.BLAS.Device.withContext' has returned is safe.
data DVector a = DVector
    { dvCtx    :: !Context
    , dvPtr    :: !(ForeignPtr (LbkVec a))
    , dvLength :: !Int
    , dvLayout :: !Layout
    }

-- | A non-zero status of the C ABI (@enum lbk_status@) with @lbk_last_error@'s text.
data LbkError = LbkError Int String deriving Show
instance Exception LbkError
