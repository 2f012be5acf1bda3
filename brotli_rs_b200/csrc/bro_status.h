// Status codes shared by the CUDA decoder, the C ABI (include/brotli_b200.h) and the tests.
// 1..24 are the reference's DecompressorError variants in enum order (src/lib.rs:294-319).
#pragma once

#define BRO_ST_OK 0
#define BRO_ST_CodeLengthsChecksum 1
#define BRO_ST_ExpectedEndOfStream 2
#define BRO_ST_ExceededExpectedBytes 3
#define BRO_ST_InvalidBlockCountCode 4
#define BRO_ST_InvalidBlockSwitchCommandCode 5
#define BRO_ST_InvalidLengthInStaticDictionary 6
#define BRO_ST_InvalidMSkipLen 7
#define BRO_ST_InvalidSymbol 8
#define BRO_ST_InvalidTransformId 9
#define BRO_ST_InvalidNonPositiveDistance 10
#define BRO_ST_LessThanTwoNonZeroCodeLengths 11
#define BRO_ST_NoCodeLength 12
#define BRO_ST_NonZeroFillBit 13
#define BRO_ST_NonZeroReservedBit 14
#define BRO_ST_NonZeroTrailerBit 15
#define BRO_ST_NonZeroTrailerNibble 16
#define BRO_ST_ParseErrorContextMap 17
#define BRO_ST_ParseErrorComplexPrefixCodeLengths 18
#define BRO_ST_ParseErrorDistanceCode 19
#define BRO_ST_ParseErrorInsertAndCopyLength 20
#define BRO_ST_ParseErrorInsertLiterals 21
#define BRO_ST_RingBufferError 22
#define BRO_ST_RunLengthExceededSizeOfContextMap 23
#define BRO_ST_UnexpectedEOF 24
/* conditions the reference cannot express as an error value */
#define BRO_ST_OutputTooSmall 100      /* caller's output slot is smaller than the decoded stream */
#define BRO_ST_CudaError 101           /* a CUDA runtime call failed (host-side entry points only) */
#define BRO_ST_PanicUppercaseZero 102  /* the reference reaches unreachable!() at src/transformation/mod.rs:78 */
#define BRO_ST_InvalidArgument 104
#define BRO_ST_SizeUnknown 103        /* bro_batch_sizes: the stream's size is only known after decoding it (literal context modelling) */
/* internal: set by the two-phase path (parse kernel) for streams it hands to the fused warp kernel's retry pass; never
 * visible to a caller */
#define BRO_ST_ArenaTooSmall 105      /* meta-block needs more table space than a parse-thread arena */
#define BRO_ST_NeedFused 106          /* literal context modelling needs the bytes of copies not yet materialised */
#define BRO_ST_RecordsFull 107        /* more copy records than the stream's share of the record arena */
#define BRO_ST_IS_RETRY(st) ((st) >= BRO_ST_ArenaTooSmall && (st) <= BRO_ST_RecordsFull)
