/* Pre-included (nvcc -include) when the reference's CUDA translation units are compiled for sm_100a by
 * oracle/build_ref_cuda.py.  The reference calls thrust algorithms whose headers it never includes (CUDA 10/11-era
 * thrust pulled them in transitively; CUDA 12.9's does not).  This file holds no code: NVIDIA headers only. */
#pragma once
#include <thrust/adjacent_difference.h>
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/count.h>
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/fill.h>
#include <thrust/for_each.h>
#include <thrust/functional.h>
#include <thrust/gather.h>
#include <thrust/host_vector.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/reduce.h>
#include <thrust/remove.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/transform.h>
#include <thrust/tuple.h>
#include <thrust/unique.h>
