// TEST INFRASTRUCTURE (oracle) -- minimal stand-in for <Eigen\SparseCore>.
//
// The reference keeps its global matrices in
//   Eigen::SparseMatrix<double, RowMajor, int>   (reference SparseMatrix.h:19)
// filled by setFromTriplets (reference SparseMatrix.cpp:67-71).  Eigen is an
// un-vendored dependency (cloned unpinned by install_dependencies.bat:22-33),
// so it is absent from /root/reference and from this image.  This header
// restates the published Eigen 3.4 algorithm of setFromTriplets for a
// row-major target:
//   1. bucket the triplets by COLUMN, preserving insertion order;
//   2. collapse duplicates inside each column, adding later values onto the
//      first occurrence (summation in insertion order);
//   3. transpose-copy into row-major storage, which leaves the column indices
//      of every row sorted ascending.  Explicit zeros are kept.
// Only the members the reference's SparseMatrix.h/.cpp and Modal/Static code
// touch are provided.
#pragma once
#include <vector>
#include <cstddef>

namespace Eigen {

template <typename Scalar, typename StorageIndex = int>
class Triplet
{
public:
	Triplet() : m_row(0), m_col(0), m_value(0) {}
	Triplet(const StorageIndex& i, const StorageIndex& j, const Scalar& v = Scalar(0))
		: m_row(i), m_col(j), m_value(v) {}
	const StorageIndex& row() const { return m_row; }
	const StorageIndex& col() const { return m_col; }
	const Scalar& value() const { return m_value; }
protected:
	StorageIndex m_row, m_col;
	Scalar m_value;
};

template <typename Scalar, int Options, typename StorageIndex>
class SparseMatrix
{
	static_assert(Options == 1, "oracle stub: only RowMajor is restated");
public:
	typedef long Index;
	SparseMatrix() : m_rows(0), m_cols(0), m_outer(1, 0) {}

	void resize(Index r, Index c)
	{
		m_rows = r; m_cols = c;
		m_outer.assign((size_t)r + 1, 0);
		m_inner.clear(); m_values.clear();
	}
	void reserve(Index n) { m_inner.reserve((size_t)n); m_values.reserve((size_t)n); }
	Index rows() const { return m_rows; }
	Index cols() const { return m_cols; }
	Index outerSize() const { return m_rows; }
	Index nonZeros() const { return (Index)m_values.size(); }
	Scalar* valuePtr() { return m_values.data(); }
	StorageIndex* innerIndexPtr() { return m_inner.data(); }
	StorageIndex* outerIndexPtr() { return m_outer.data(); }
	const Scalar* valuePtr() const { return m_values.data(); }
	const StorageIndex* innerIndexPtr() const { return m_inner.data(); }
	const StorageIndex* outerIndexPtr() const { return m_outer.data(); }

	template <typename It>
	void setFromTriplets(It begin, It end)
	{
		const size_t nr = (size_t)m_rows, nc = (size_t)m_cols;
		// pass 1: column buckets in insertion order
		std::vector<size_t> cstart(nc + 1, 0);
		for (It t = begin; t != end; ++t) cstart[(size_t)t->col() + 1]++;
		for (size_t j = 0; j < nc; j++) cstart[j + 1] += cstart[j];
		const size_t nt = cstart[nc];
		std::vector<StorageIndex> brow(nt);
		std::vector<Scalar> bval(nt);
		{
			std::vector<size_t> fill(cstart.begin(), cstart.end() - 1);
			for (It t = begin; t != end; ++t)
			{
				size_t p = fill[(size_t)t->col()]++;
				brow[p] = t->row();
				bval[p] = t->value();
			}
		}
		// pass 2: collapse duplicates per column (first occurrence keeps the slot)
		std::vector<long> seen(nr, -1);
		std::vector<size_t> cend(nc, 0);
		size_t w = 0;
		std::vector<size_t> cbeg(nc, 0);
		for (size_t j = 0; j < nc; j++)
		{
			const size_t start = w;
			cbeg[j] = start;
			for (size_t p = cstart[j]; p < cstart[j + 1]; p++)
			{
				const size_t i = (size_t)brow[p];
				if (seen[i] >= (long)start)
					bval[(size_t)seen[i]] += bval[p];
				else
				{
					brow[w] = brow[p];
					bval[w] = bval[p];
					seen[i] = (long)w;
					w++;
				}
			}
			cend[j] = w;
		}
		// pass 3: transpose-copy to row-major (columns ascending inside a row)
		m_outer.assign(nr + 1, 0);
		for (size_t p = 0; p < w; p++) m_outer[(size_t)brow[p] + 1]++;
		for (size_t i = 0; i < nr; i++) m_outer[i + 1] += m_outer[i];
		m_inner.resize(w);
		m_values.resize(w);
		std::vector<StorageIndex> rfill(m_outer.begin(), m_outer.end() - 1);
		for (size_t j = 0; j < nc; j++)
			for (size_t p = cbeg[j]; p < cend[j]; p++)
			{
				const StorageIndex q = rfill[(size_t)brow[p]]++;
				m_inner[(size_t)q] = (StorageIndex)j;
				m_values[(size_t)q] = bval[p];
			}
	}

	class InnerIterator
	{
	public:
		InnerIterator(const SparseMatrix& m, Index outer)
			: m_m(m), m_outerIdx(outer), m_p(m.m_outer[(size_t)outer]), m_e(m.m_outer[(size_t)outer + 1]) {}
		operator bool() const { return m_p < m_e; }
		InnerIterator& operator++() { ++m_p; return *this; }
		Index row() const { return m_outerIdx; }
		Index col() const { return m_m.m_inner[(size_t)m_p]; }
		Scalar value() const { return m_m.m_values[(size_t)m_p]; }
	private:
		const SparseMatrix& m_m;
		Index m_outerIdx;
		StorageIndex m_p, m_e;
	};

private:
	Index m_rows, m_cols;
	std::vector<StorageIndex> m_outer;
	std::vector<StorageIndex> m_inner;
	std::vector<Scalar> m_values;
};

} // namespace Eigen
