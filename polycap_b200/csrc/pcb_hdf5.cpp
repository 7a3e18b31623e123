/*
 * pcb_hdf5.cpp -- polycap_transmission_efficiencies_write_hdf5 without libhdf5.
 *
 * The reference writes its result file through the HDF5 C library
 * (/root/reference/src/polycap-transmission-efficiencies.c:318-780); that library is a third-party dependency which is
 * absent from this image.  This file restates the part of the published HDF5 File Format Specification (version 1.1 of
 * the format: superblock version 0, version-1 object headers, old-style groups) that the reference's file needs:
 * groups, contiguous little-endian IEEE-754 double datasets of rank 1 or 2, and one scalar fixed-length string attribute
 * ("Units") per dataset.  Same dataset names, shapes, contents and units as the reference.
 *
 *   offset 0     superblock (96 bytes, root symbol-table entry included)
 *   then         per group: object header -> B-tree node -> symbol-table node -> local heap (+ names)
 *   then         per dataset: object header (dataspace, datatype, fill value, contiguous layout, attribute)
 *   then         the raw data, 8-byte aligned, in the order of creation
 *
 * NOT validated against libhdf5/h5py here (neither exists in this image); the layout follows the specification and
 * tests/test_hdf5.py re-reads the file with an independent minimal parser.
 */
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "pcb_internal.h"

namespace {

const uint64_t UNDEF = ~0ULL;
const int GROUP_LEAF_K = 32;     /* a symbol-table node holds 2K = 64 entries: one node per group is enough here */
const int GROUP_INTERNAL_K = 16;

struct Dset {
	std::string name, units;
	int rank = 1;
	uint64_t dims[2] = {0, 0};
	const double *data = nullptr;
	std::vector<double> owned;
	uint64_t ohdr = 0, data_addr = 0;
	uint64_t count() const { return rank == 1 ? dims[0] : dims[0] * dims[1]; }
};

struct Group {
	std::string name;
	std::vector<std::unique_ptr<Group>> groups;
	std::vector<std::unique_ptr<Dset>> dsets;
	uint64_t ohdr = 0, btree = 0, snod = 0, heap = 0, heap_data = 0, heap_size = 0;
	std::vector<std::pair<std::string, uint64_t>> name_off; /* member name -> offset in the heap data segment */
};

uint64_t pad8(uint64_t n) { return (n + 7) & ~7ULL; }

struct Buf {
	std::vector<uint8_t> b;
	void u8(uint8_t v) { b.push_back(v); }
	void u16(uint16_t v) { for (int i = 0; i < 2; i++) b.push_back((uint8_t)(v >> (8 * i))); }
	void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
	void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
	void bytes(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
	void zeros(size_t n) { b.insert(b.end(), n, 0); }
	void pad_to(size_t n) { if (b.size() < n) zeros(n - b.size()); }
	void align8() { zeros((size_t)(pad8(b.size()) - b.size())); }
};

/* sizes of the metadata blocks */
const uint64_t SUPERBLOCK_SIZE = 96;
const uint64_t GROUP_OHDR_SIZE = 16 + 8 + 16;
const uint64_t BTREE_SIZE = 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8;
const uint64_t SNOD_SIZE = 8 + 2 * GROUP_LEAF_K * 40;
const uint64_t HEAP_HDR_SIZE = 32;

uint64_t attr_msg_size(const Dset &d)
{
	/* attribute v1: 8 header + name "Units\0" padded to 8 + string datatype (8) + scalar dataspace (8) + data */
	return pad8(8 + 8 + 8 + 8 + d.units.size());
}

uint64_t dset_ohdr_size(const Dset &d)
{
	uint64_t msgs = (8 + pad8(8 + 8 * (uint64_t)d.rank)) /* dataspace */ + (8 + 24) /* datatype */ + (8 + 8) /* fill value */ + (8 + 24) /* layout */ +
			(8 + attr_msg_size(d));
	return 16 + msgs;
}

void layout(Group &g, uint64_t &pos)
{
	g.ohdr = pos;
	pos += GROUP_OHDR_SIZE;
	g.btree = pos;
	pos += BTREE_SIZE;
	g.snod = pos;
	pos += SNOD_SIZE;
	g.heap = pos;
	pos += HEAP_HDR_SIZE;
	/* heap data: the empty string at offset 0, then the member names, each NUL-terminated and padded to 8; a free block closes it */
	std::vector<std::string> names;
	for (auto &c : g.groups) names.push_back(c->name);
	for (auto &d : g.dsets) names.push_back(d->name);
	std::sort(names.begin(), names.end());
	uint64_t off = 8;
	g.name_off.clear();
	for (auto &n : names) {
		g.name_off.emplace_back(n, off);
		off += pad8(n.size() + 1);
	}
	g.heap_size = off + 16; /* + one free block (next, size) */
	g.heap_data = pos;
	pos += g.heap_size;
	for (auto &c : g.groups)
		layout(*c, pos);
	for (auto &d : g.dsets) {
		d->ohdr = pos;
		pos += dset_ohdr_size(*d);
	}
}

void layout_data(Group &g, uint64_t &pos)
{
	for (auto &c : g.groups)
		layout_data(*c, pos);
	for (auto &d : g.dsets) {
		d->data_addr = pos;
		pos += pad8(d->count() * 8);
	}
}

uint64_t heap_offset(const Group &g, const std::string &n)
{
	for (auto &p : g.name_off)
		if (p.first == n)
			return p.second;
	return 0;
}

void msg_header(Buf &b, uint16_t type, uint16_t size)
{
	b.u16(type);
	b.u16(size);
	b.u8(0); /* flags */
	b.zeros(3);
}

void emit_group(const Group &g, Buf &b)
{
	/* ---- object header (version 1) with one Symbol Table message ---- */
	b.pad_to((size_t)g.ohdr);
	b.u8(1); b.u8(0); b.u16(1); /* version, reserved, number of messages */
	b.u32(1);                   /* object reference count */
	b.u32(8 + 16);              /* header size: bytes of messages */
	b.zeros(4);                 /* alignment of the first message */
	msg_header(b, 0x0011, 16);
	b.u64(g.btree);
	b.u64(g.heap);
	/* ---- B-tree node, type 0 (group), leaf level, one child ---- */
	b.pad_to((size_t)g.btree);
	b.bytes("TREE", 4);
	b.u8(0); b.u8(0);
	const bool has_members = !g.name_off.empty();
	b.u16(has_members ? 1 : 0); /* entries used */
	b.u64(UNDEF); b.u64(UNDEF); /* siblings */
	b.u64(0);                   /* key 0: the empty string */
	if (has_members) {
		b.u64(g.snod);
		b.u64(g.name_off.back().second); /* key 1: the largest name of the child */
	}
	b.pad_to((size_t)(g.btree + BTREE_SIZE));
	/* ---- symbol-table node: entries sorted by name ---- */
	b.pad_to((size_t)g.snod);
	b.bytes("SNOD", 4);
	b.u8(1); b.u8(0);
	b.u16((uint16_t)g.name_off.size());
	for (auto &p : g.name_off) {
		const Group *cg = nullptr;
		const Dset *cd = nullptr;
		for (auto &c : g.groups) if (c->name == p.first) cg = c.get();
		for (auto &d : g.dsets) if (d->name == p.first) cd = d.get();
		b.u64(p.second);                  /* link name offset in the heap */
		b.u64(cg ? cg->ohdr : cd->ohdr);  /* object header address */
		b.u32(cg ? 1 : 0);                /* cache type: 1 = group (scratch pad holds B-tree and heap addresses) */
		b.u32(0);
		if (cg) {
			b.u64(cg->btree);
			b.u64(cg->heap);
		} else {
			b.zeros(16);
		}
	}
	b.pad_to((size_t)(g.snod + SNOD_SIZE));
	/* ---- local heap ---- */
	b.pad_to((size_t)g.heap);
	b.bytes("HEAP", 4);
	b.u8(0); b.zeros(3);
	b.u64(g.heap_size);
	b.u64(g.heap_size - 16); /* head of the free list: the closing free block */
	b.u64(g.heap_data);
	b.pad_to((size_t)g.heap_data);
	b.zeros(8); /* offset 0: "" */
	for (auto &p : g.name_off) {
		b.pad_to((size_t)(g.heap_data + p.second));
		b.bytes(p.first.c_str(), p.first.size() + 1);
	}
	b.pad_to((size_t)(g.heap_data + g.heap_size - 16));
	b.u64(1);  /* next free block: H5HL_FREE_NULL */
	b.u64(16); /* size of this free block */
	for (auto &c : g.groups)
		emit_group(*c, b);
	for (auto &dp : g.dsets) {
		const Dset &d = *dp;
		b.pad_to((size_t)d.ohdr);
		const uint64_t total = dset_ohdr_size(d);
		b.u8(1); b.u8(0); b.u16(5);
		b.u32(1);
		b.u32((uint32_t)(total - 16));
		b.zeros(4);
		/* dataspace, version 1 */
		const uint16_t ds = (uint16_t)pad8(8 + 8 * (uint64_t)d.rank);
		msg_header(b, 0x0001, ds);
		b.u8(1); b.u8((uint8_t)d.rank); b.u8(0); b.u8(0); b.zeros(4);
		for (int k = 0; k < d.rank; k++)
			b.u64(d.dims[k]);
		/* datatype: IEEE-754 binary64, little endian */
		msg_header(b, 0x0003, 24);
		b.u8(0x11); b.u8(0x20); b.u8(0x3f); b.u8(0x00);
		b.u32(8);
		b.u16(0); b.u16(64); b.u8(52); b.u8(11); b.u8(0); b.u8(52); b.u32(1023);
		b.zeros(4);
		/* fill value, version 2: late allocation, written if set, default fill value */
		msg_header(b, 0x0005, 8);
		b.u8(2); b.u8(2); b.u8(2); b.u8(1); b.u32(0);
		/* data layout, version 3, contiguous */
		msg_header(b, 0x0008, 24);
		b.u8(3); b.u8(1);
		b.u64(d.data_addr);
		b.u64(d.count() * 8);
		b.zeros(6);
		/* attribute "Units": scalar fixed-length string */
		const uint64_t asz = attr_msg_size(d);
		msg_header(b, 0x000C, (uint16_t)asz);
		const size_t a0 = b.b.size();
		b.u8(1); b.u8(0);
		b.u16(6);                           /* name size incl. NUL */
		b.u16(8);                           /* datatype size */
		b.u16(8);                           /* dataspace size */
		b.bytes("Units\0\0\0", 8);
		b.u8(0x13); b.u8(0x00); b.u8(0); b.u8(0); /* string, null-terminated, ASCII */
		b.u32((uint32_t)d.units.size());
		b.u8(1); b.u8(0); b.u8(0); b.u8(0); b.zeros(4); /* scalar dataspace */
		b.bytes(d.units.data(), d.units.size());
		b.pad_to(a0 + (size_t)asz);
	}
}

Dset *add(Group &g, const char *name, const char *units, int rank, uint64_t d0, uint64_t d1, const double *data)
{
	std::unique_ptr<Dset> d(new Dset());
	d->name = name;
	d->units = units;
	d->rank = rank;
	d->dims[0] = d0;
	d->dims[1] = d1;
	d->data = data;
	g.dsets.push_back(std::move(d));
	return g.dsets.back().get();
}

/* [k][n] row-major from k component arrays of length n (the reference's data_temp packing) */
Dset *add_rows(Group &g, const char *name, const char *units, int k, uint64_t n, double *const *rows)
{
	Dset *d = add(g, name, units, 2, (uint64_t)k, n, nullptr);
	d->owned.resize((size_t)k * n);
	for (int r = 0; r < k; r++)
		memcpy(d->owned.data() + (size_t)r * n, rows[r], n * 8);
	d->data = d->owned.data();
	return d;
}

Dset *add_i64(Group &g, const char *name, const char *units, uint64_t n, const int64_t *v)
{
	Dset *d = add(g, name, units, 1, n, 0, nullptr);
	d->owned.resize(n);
	for (uint64_t i = 0; i < n; i++)
		d->owned[i] = (double)v[i];
	d->data = d->owned.data();
	return d;
}

Group *subgroup(Group &g, const char *name)
{
	std::unique_ptr<Group> c(new Group());
	c->name = name;
	g.groups.push_back(std::move(c));
	return g.groups.back().get();
}

bool write_data(const Group &g, FILE *fp, uint64_t &pos)
{
	static const uint8_t zero[8] = {0};
	for (auto &c : g.groups)
		if (!write_data(*c, fp, pos))
			return false;
	for (auto &d : g.dsets) {
		const uint64_t n = d->count() * 8;
		if (n && fwrite(d->data, 1, n, fp) != n)
			return false;
		const uint64_t padn = pad8(n) - n;
		if (padn && fwrite(zero, 1, padn, fp) != padn)
			return false;
		pos += pad8(n);
	}
	return true;
}

} // namespace

extern "C" int pcb_hdf5_write(polycap_transmission_efficiencies *eff, const char *filename, char *errbuf, size_t errlen)
{
	struct _polycap_images *im = eff->images;
	polycap_source *src = eff->source;
	if (!im || !src || !src->description || !src->description->profile) {
		snprintf(errbuf, errlen, "efficiencies carry no images/source");
		return -1;
	}
	const uint64_t nE = eff->n_energies, nx = (uint64_t)im->i_exit;
	polycap_description *de = src->description;
	polycap_profile *pr = de->profile;

	Group root;
	/* the reference's datasets, names, shapes and units: transmission-efficiencies.c:345-766 */
	add(root, "Energies", "keV", 1, nE, 0, eff->energies);
	add(root, "Transmission_Efficiencies", "a.u.", 1, nE, 0, eff->efficiencies);
	add_rows(root, "Source_Start_Coordinates", "[cm,cm]", 2, nx, im->src_start_coords);
	Group *st = subgroup(root, "PC_Start");
	add_rows(*st, "Coordinates", "[cm,cm]", 2, nx, im->pc_start_coords);
	add_rows(*st, "Direction", "[cm,cm]", 2, nx, im->pc_start_dir);
	add_rows(*st, "Electric_Vector", "[cm,cm]", 2, nx, im->pc_start_elecv);
	Group *ex = subgroup(root, "PC_Exit");
	add_rows(*ex, "Coordinates", "[cm,cm,cm]", 3, nx, im->pc_exit_coords);
	add_i64(*ex, "N_Reflections", "a.u.", nx, im->pc_exit_nrefl);
	add_rows(*ex, "Direction", "[cm,cm]", 2, nx, im->pc_exit_dir);
	add_rows(*ex, "Electric_Vector", "[cm,cm]", 2, nx, im->pc_exit_elecv);
	add(*ex, "Weights", "[keV,a.u.]", 2, nx, nE, im->exit_coord_weights);
	add(*ex, "D_Travel", "[cm]", 1, nx, 0, im->pc_exit_dtravel);
	for (int kind = 0; kind < 2; kind++) {
		const uint64_t nl = (uint64_t)(kind == 0 ? im->i_extleak : im->i_intleak);
		if (nl == 0)
			continue;
		Group *lg = subgroup(root, kind == 0 ? "ExternalLeaks" : "InternalLeaks");
		double *const *co = kind == 0 ? im->extleak_coords : im->intleak_coords;
		double *const *di = kind == 0 ? im->extleak_dir : im->intleak_dir;
		const double *w = kind == 0 ? im->extleak_coord_weights : im->intleak_coord_weights;
		add_rows(*lg, "Coordinates", "[cm,cm,cm]", 3, nl, co);
		add_rows(*lg, "Direction", "[cm,cm]", 2, nl, di);
		if (kind == 1)
			add_rows(*lg, "Electric_Vector", "[cm,cm]", 2, nl, im->intleak_elecv);
		add(*lg, "Weights", "[keV,a.u.]", 2, nl, nE, w);
		Dset *tot = add(*lg, "Weight_Total", "a.u.", 1, nE, 0, nullptr);
		tot->owned.assign(nE, 0.0);
		for (uint64_t j = 0; j < nE; j++) {
			for (uint64_t k = 0; k < nl; k++)
				tot->owned[j] += w[k * nE + j];
			tot->owned[j] = tot->owned[j] / (double)im->i_start;
		}
		tot->data = tot->owned.data();
		add_i64(*lg, "N_Reflections", "a.u.", nl, kind == 0 ? im->extleak_n_refl : im->intleak_n_refl);
	}
	Group *in = subgroup(root, "Input");
	{
		const uint64_t nm = (uint64_t)pr->nmax;
		double *rows[2] = {pr->z, pr->ext};
		add_rows(*in, "PC_Shape", "[cm,cm]", 2, nm, rows);
		double *rows2[2] = {pr->z, pr->cap};
		add_rows(*in, "Cap_Shape", "[cm,cm]", 2, nm, rows2);
		Dset *nc = add(*in, "N_Capillaries", "a.u.", 1, 1, 0, nullptr);
		nc->owned.assign(1, (double)de->n_cap);
		nc->data = nc->owned.data();
		add(*in, "Surface_Roughness", "Angstrom", 1, 1, 0, &de->sig_rough);
		add(*in, "Open_Area", "a.u.", 1, 1, 0, &de->open_area);
		Dset *comp = add(*in, "PC_Composition", "[Z,w%]", 2, 2, de->nelem, nullptr);
		comp->owned.resize(2 * (size_t)de->nelem);
		for (unsigned j = 0; j < de->nelem; j++) {
			comp->owned[j] = de->iz[j];
			comp->owned[j + de->nelem] = de->wi[j];
		}
		comp->data = comp->owned.data();
		add(*in, "PC_Density", "g/cm3", 1, 1, 0, &de->density);
		add(*in, "Src_PC_Dist", "cm", 1, 1, 0, &src->d_source);
	}

	uint64_t pos = SUPERBLOCK_SIZE;
	layout(root, pos);
	const uint64_t meta_end = pad8(pos);
	uint64_t dpos = meta_end;
	layout_data(root, dpos);
	const uint64_t eof = dpos;

	Buf b;
	b.bytes("\x89HDF\r\n\x1a\n", 8);
	b.u8(0); b.u8(0); b.u8(0); b.u8(0); /* superblock, free-space, root-group versions; reserved */
	b.u8(0);                            /* shared header message format version */
	b.u8(8); b.u8(8); b.u8(0);          /* size of offsets, size of lengths, reserved */
	b.u16(GROUP_LEAF_K);
	b.u16(GROUP_INTERNAL_K);
	b.u32(0);                           /* file consistency flags */
	b.u64(0);                           /* base address */
	b.u64(UNDEF);                       /* free-space info */
	b.u64(eof);                         /* end of file */
	b.u64(UNDEF);                       /* driver information block */
	b.u64(0);                           /* root entry: link name offset */
	b.u64(root.ohdr);
	b.u32(1); b.u32(0);
	b.u64(root.btree);
	b.u64(root.heap);
	emit_group(root, b);
	b.pad_to((size_t)meta_end);

	FILE *fp = fopen(filename, "wb");
	if (!fp) {
		snprintf(errbuf, errlen, "could not open %s for writing -> %s", filename, strerror(errno));
		return -1;
	}
	bool ok = fwrite(b.b.data(), 1, b.b.size(), fp) == b.b.size();
	uint64_t wpos = meta_end;
	ok = ok && write_data(root, fp, wpos) && wpos == eof;
	if (fclose(fp) != 0)
		ok = false;
	if (!ok) {
		snprintf(errbuf, errlen, "write to %s failed -> %s", filename, strerror(errno));
		return -1;
	}
	return 0;
}
