//! `lapack::laswp` stays host-side: it is generic over any `T: Copy` (it is also applied to the
//! `usize` permutation vector in decomposition/lu.rs) and is pure index work.  Same contract as
//! the reference (src/lapack/laswp.rs:11-40).
use std::ptr;

#[allow(clippy::cast_possible_wrap)]
pub unsafe fn laswp<T>(ncols: usize, a: *mut T, row_stride: isize, col_stride: isize, begin: usize, piv: &[usize])
where
    T: Copy,
{
    for (i, &p) in piv.iter().enumerate().skip(begin) {
        if i == p {
            continue;
        }
        let mut row1 = a.offset(i as isize * row_stride);
        let mut row2 = a.offset(p as isize * row_stride);
        for _ in 0..ncols {
            ptr::swap(row1, row2);
            row1 = row1.offset(col_stride);
            row2 = row2.offset(col_stride);
        }
    }
}
