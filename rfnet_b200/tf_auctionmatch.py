"""Drop-in for the reference's tf_ops/emd/tf_auctionmatch.py."""
from . import ops


def auction_match(xyz1, xyz2):
    '''
input:
	xyz1 : batch_size * #points * 3
	xyz2 : batch_size * #points * 3
returns:
	matchl : batch_size * #npoints     index into xyz2 assigned to every point of xyz1
	matchr : batch_size * #npoints     index into xyz1 assigned to every point of xyz2

No gradient (ops.NoGradient('AuctionMatch'), tf_auctionmatch.py:21).
    '''
    return ops.auction_match_op(xyz1.detach(), xyz2.detach())
