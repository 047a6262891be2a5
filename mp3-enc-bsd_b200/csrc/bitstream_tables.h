// bitstream_tables.h — constant tables of the device bitstream formatter (bitstream.cuh), host-built (tables.cpp)
#pragma once
namespace mp3gpu {
struct BitTables {
    unsigned int hcode[1412];       // Table B.7: code | len << 24, flat, tables 0..33
    unsigned short hoff[34];
    unsigned char ylen[34], linbits[34];
    unsigned short short_e0[288];   // short blocks: emission pair p -> index of x in ix[576]; y = x + 3 (l3bitstream.c:546-575)
    short sfb_l[24];
    unsigned char header[4];        // the 32 header bits (constant per stream format), l3bitstream.c:323-336
    int frame_bytes, si_bytes, n_ch, pad;
};

}  // namespace mp3gpu
