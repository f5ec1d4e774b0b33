{- | Properties for "Numerical.BLAS.Device" in the style of the reference's own suite (test/Main.hs:28-52): where that suite
states @native == Fortran@ over random inputs, this one states @device == native@ over the same domains -- same generators
('Gen.genNiceFloat', 'Gen.genNVector'), same n / increment / length ranges, same 'runTest' reporting.  Elementwise routines
and @iamax@ must agree exactly; the reductions of the device path add in a different order (a tree instead of a left fold),
so @dot@ / @asum@ / @nrm2@ are held to the bound @2 n eps sum|x_i y_i|@ (DESIGN.md, section 1).

To run it, a maintainer adds this file next to test/Main.hs as a second test-suite (hs/patches/lambda-blas.cabal.patch lists
it: @main-is: DeviceMain.hs@, @other-modules: Gen@, @extra-libraries: lbk@) on a machine with a B200 and @liblbk.so@.

NOTE: no GHC in this image: delivered as source, linted by tests/test_haskell_shim.py, never compiled here.
-}
module Main( main ) where

import qualified Data.Vector.Storable as V
import Test.Tasty
import Test.Tasty.QuickCheck
-- | The reference
import Numerical.BLAS.Single
-- | The device path
import qualified Numerical.BLAS.Device as D
-- | The reference's generators (test/Gen.hs)
import Gen

main :: IO ()
main = D.withContext 0 $ \ ctx -> defaultMain (tests ctx)

tests :: D.Context -> TestTree
tests ctx = testGroup "BLAS on the device"
    [ testGroup "Level-1"
        [ dotTest ctx "sdot" genNiceFloat (elements [-5..5])
        , iviTest ctx "sasum" asum D.asum genNiceFloat (elements [1..5])
        , iviTest ctx "snrm2" nrm2 D.nrm2 genNiceFloat (elements [1..5])
        , iamaxTest ctx "isamax" genNiceFloat (elements [1..5])
        , scalTest ctx "sscal" genNiceFloat (elements [1..5])
        , copyTest ctx "scopy" genNiceFloat (elements [-5..5])
        , swapTest ctx "sswap" genNiceFloat (elements [-5..5] `suchThat` (/=0))
        , rotTest  ctx "srot"  genNiceFloat (elements [-5,5])
        , axpyTest ctx "saxpy" genNiceFloat (elements [-5,5])
        ]
    ]

-- | The device's sum is within the reduction bound of the reference's left fold.
within :: Float -> Float -> Float -> Bool
within bound expected observed = abs (expected - observed) <= bound

-- | cf. test/Main.hs:59-79
dotTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
dotTest ctx testname gen genInc = testProperty testname $
    forAll (choose (1,100)) $ \ n ->
    forAll genInc $ \ incx ->
    forAll genInc $ \ incy ->
    forAll (genNVector gen (1+(n-1)*abs incx )) $ \ u ->
    forAll (genNVector gen (1+(n-1)*abs incy )) $ \ v ->
       ioProperty $ do
           du <- D.toDevice ctx u
           dv <- D.toDevice ctx v
           observed <- D.dot n du incx dv incy
           let expected = dot n u incx v incy
               bound = 2 * fromIntegral n * epsilonF * asum n (V.zipWith (*) (spread n u incx) (spread n v incy)) 1
           reportWithin bound expected observed

-- | cf. test/Main.hs:239-257 (asum, nrm2)
iviTest :: D.Context -> String
    -> (Int -> V.Vector Float -> Int -> Float)
    -> (Int -> D.DVector Float -> Int -> IO Float)
    -> Gen Float -> Gen Int -> TestTree
iviTest ctx testname func dfunc gen genInc = testProperty testname $
    forAll (choose (0,100)) $ \ n ->
    forAll genInc $ \ incx ->
    forAll (genNVector gen (max 1 (1+(n-1)*incx))) $ \ u ->
       ioProperty $ do
           du <- D.toDevice ctx u
           observed <- dfunc n du incx
           let expected = func n u incx
               bound = 2 * fromIntegral n * epsilonF * max expected (asum n u incx)
           reportWithin bound expected observed

-- | cf. test/Main.hs:41 -- the index must be the same one
iamaxTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
iamaxTest ctx testname gen genInc = testProperty testname $
    forAll (choose (0,100)) $ \ n ->
    forAll genInc $ \ incx ->
    forAll (genNVector gen (max 1 (1+(n-1)*incx))) $ \ u ->
       ioProperty $ do
           du <- D.toDevice ctx u
           observed <- D.iamax n du incx
           runTest (iamax n u incx) observed

-- | cf. test/Main.hs:405-424
scalTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
scalTest ctx testname gen genInc = testProperty testname $
    forAll (choose (1,100)) $ \ n ->
    forAll gen $ \ a ->
    forAll genInc $ \ incx ->
    forAll (genNVector gen (1+(n-1)*incx )) $ \ u ->
       ioProperty $ do
           observed <- D.toDevice ctx u >>= \ du -> D.scal n a du incx >>= D.fromDevice
           runTest (scal n a u incx) observed

-- | cf. test/Main.hs:429-451
copyTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
copyTest ctx testname gen genInc = testProperty testname $
    forAll (choose (1,5)) $ \ n ->
    forAll (choose (1,2)) $ \ mx ->
    forAll (choose (1,2)) $ \ my ->
    forAll genInc $ \ incx ->
    forAll genInc $ \ incy ->
    forAll (genNVector gen (mx+1+(n-1)*abs incx )) $ \ u ->
    forAll (genNVector gen (my+1+(n-1)*abs incy )) $ \ v ->
       ioProperty $ do
           du <- D.toDevice ctx u
           dv <- D.toDevice ctx v
           observed <- D.copy n du incx dv incy >>= D.fromDevice
           runTest (copy n u incx v incy) observed

-- | cf. test/Main.hs:107-129
swapTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
swapTest ctx testname gen genInc = testProperty testname $
   forAll (choose (1,5)) $ \ n ->
   forAll (choose (0,2)) $ \ mx ->
   forAll (choose (0,2)) $ \ my ->
   forAll genInc $ \ incx ->
   forAll genInc $ \ incy ->
   forAll (genNVector gen (mx+1+(n-1)*abs incx )) $ \ u ->
   forAll (genNVector gen (my+1+(n-1)*abs incy )) $ \ v ->
      ioProperty $ do
          du <- D.toDevice ctx u
          dv <- D.toDevice ctx v
          (dx,dy) <- D.swap n du incx dv incy
          x <- D.fromDevice dx
          y <- D.fromDevice dy
          runTest (swap n u incx v incy) (x,y)

-- | cf. test/Main.hs:153-176
rotTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
rotTest ctx testname gen genInc = testProperty testname $
   forAll (choose (1,100)) $ \ n ->
   forAll gen $ \ a ->
   forAll genInc $ \ incx ->
   forAll genInc $ \ incy ->
   forAll (genNVector gen (1+(n-1)*abs incx )) $ \ u ->
   forAll (genNVector gen (1+(n-1)*abs incy )) $ \ v ->
      ioProperty $ do
          let c = cos a
              s = sin a
          du <- D.toDevice ctx u
          dv <- D.toDevice ctx v
          (dx,dy) <- D.rot n du incx dv incy c s
          x <- D.fromDevice dx
          y <- D.fromDevice dy
          runTest (rot n u incx v incy c s) (x,y)

-- | cf. test/Main.hs:456-477
axpyTest :: D.Context -> String -> Gen Float -> Gen Int -> TestTree
axpyTest ctx testname gen genInc = testProperty testname $
    forAll (choose (1,100)) $ \ n ->
    forAll gen $ \ a ->
    forAll genInc $ \ incx ->
    forAll genInc $ \ incy ->
    forAll (choose (1,100)) $ \ mx ->
    forAll (choose (1,100)) $ \ my ->
    forAll (genNVector gen (mx+(n-1)*abs incx )) $ \ u ->
    forAll (genNVector gen (my+(n-1)*abs incy )) $ \ v ->
       ioProperty $ do
           du <- D.toDevice ctx u
           dv <- D.toDevice ctx v
           observed <- D.axpy n a du incx dv incy >>= D.fromDevice
           runTest (axpy n a u incx v incy) observed

-- | The n elements a routine visits, as a dense vector (for the bound of 'dotTest').
spread :: Int -> V.Vector Float -> Int -> V.Vector Float
spread n v inc = V.generate n pick
    where
    first = if inc > 0 then 0 else (1-n)*inc
    pick i = V.unsafeIndex v (first + i*inc)

-- | Unit round-off of Float.
epsilonF :: Float
epsilonF = 5.9604645e-8

-- | 'runTest' of test/Main.hs:262-270 for a result that is only bounded.
reportWithin :: Float -> Float -> Float -> IO Bool
reportWithin bound expected observed =
  if within bound expected observed
        then return True
        else do
            putStrLn ""
            putStrLn $ "EXPECTED: " ++ show expected ++ " +/- " ++ show bound
            putStrLn $ "OBSERVED: " ++ show observed
            return False

-- | test/Main.hs:262-270
runTest :: (Show a, Eq a) => a -> a -> IO Bool
runTest expected observed =
  if expected==observed
        then return True
        else do
            putStrLn ""
            putStrLn $ "EXPECTED: " ++ show expected
            putStrLn $ "OBSERVED: " ++ show observed
            return False
