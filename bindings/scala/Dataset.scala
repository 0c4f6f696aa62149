/*
 * Scala facade a MatRel maintainer adds next to
 * src/main/scala/org/apache/spark/sql/matfast/Dataset.scala to route the block-multiply path to the
 * B200 engine.  Same method names / argument order as the reference's Dataset (Dataset.scala:57-152);
 * each call forwards to the C ABI (include/matrel.h) through the JNI shim (bindings/jni/matrel_jni.cpp).
 * NOT compiled in this image (no scalac / JDK).
 */
package org.apache.spark.sql.matfast.b200

import org.apache.spark.sql.matfast.matrix.{DenseMatrix, MLMatrix, MatrixBlock, SparseMatrix}

private[b200] object Native {
  System.loadLibrary("matrel_jni") // links libmatrel_b200.so
  @native def init(device: Int, compatBugs: Boolean): Long
  @native def shutdown(ctx: Long): Unit
  @native def matrixCreate(ctx: Long): Long
  @native def matrixFree(m: Long): Unit
  @native def putBlock(m: Long, rid: Int, cid: Int, tpe: Byte, numRows: Int, numCols: Int,
                       colPtrs: Array[Int], rowIndices: Array[Int], values: Array[Double],
                       isTransposed: Boolean): Unit
  @native def matrixMultiply(l: Long, lr: Long, lc: Long, r: Long, rr: Long, rc: Long, blk: Int): Long
  @native def addElement(l: Long, lr: Long, lc: Long, r: Long, rr: Long, rc: Long, blk: Int): Long
  @native def multiplyElement(l: Long, lr: Long, lc: Long, r: Long, rr: Long, rc: Long, blk: Int): Long
  @native def divideElement(l: Long, lr: Long, lc: Long, r: Long, rr: Long, rc: Long, blk: Int): Long
  @native def matrixRankOneUpdate(l: Long, lr: Long, lc: Long, r: Long, rr: Long, rc: Long, blk: Int): Long
  @native def transpose(a: Long): Long
  @native def addScalar(a: Long, alpha: Double): Long
  @native def multiplyScalar(a: Long, alpha: Double): Long
  @native def power(a: Long, alpha: Double): Long
}

class B200Session(device: Int = -1, compatBugs: Boolean = true) {
  private[b200] val ctx: Long = Native.init(device, compatBugs)
  def stop(): Unit = Native.shutdown(ctx)

  /** Seq(MatrixBlock(...)).toDS() (example/BasicMatrixOps.scala:115-116) */
  def toDS(blocks: Seq[MatrixBlock]): B200Dataset = {
    val h = Native.matrixCreate(ctx)
    blocks.foreach { b =>
      b.matrix match { // the 7 fields of MLMatrixSerializer.serialize (util/MLMatrixSerializer.scala:26-48)
        case d: DenseMatrix =>
          Native.putBlock(h, b.rid, b.cid, 1, d.numRows, d.numCols, null, null, d.values, d.isTransposed)
        case s: SparseMatrix =>
          Native.putBlock(h, b.rid, b.cid, 0, s.numRows, s.numCols, s.colPtrs, s.rowIndices, s.values, s.isTransposed)
      }
    }
    new B200Dataset(this, h)
  }
}

class B200Dataset private[b200](val session: B200Session, private[b200] val h: Long) {
  private def wrap(x: Long) = new B200Dataset(session, x)
  def matrixMultiply(leftRowNum: Long, leftColNum: Long, right: B200Dataset,
                     rightRowNum: Long, rightColNum: Long, blkSize: Int): B200Dataset =
    wrap(Native.matrixMultiply(h, leftRowNum, leftColNum, right.h, rightRowNum, rightColNum, blkSize))
  def addElement(leftRowNum: Long, leftColNum: Long, right: B200Dataset,
                 rightRowNum: Long, rightColNum: Long, blkSize: Int): B200Dataset =
    wrap(Native.addElement(h, leftRowNum, leftColNum, right.h, rightRowNum, rightColNum, blkSize))
  def multiplyElement(leftRowNum: Long, leftColNum: Long, right: B200Dataset,
                      rightRowNum: Long, rightColNum: Long, blkSize: Int): B200Dataset =
    wrap(Native.multiplyElement(h, leftRowNum, leftColNum, right.h, rightRowNum, rightColNum, blkSize))
  def divideElement(leftRowNum: Long, leftColNum: Long, right: B200Dataset,
                    rightRowNum: Long, rightColNum: Long, blkSize: Int): B200Dataset =
    wrap(Native.divideElement(h, leftRowNum, leftColNum, right.h, rightRowNum, rightColNum, blkSize))
  def matrixRankOneUpdate(leftRowNum: Long, leftColNum: Long, right: B200Dataset,
                          rightRowNum: Long, rightColNum: Long, blkSize: Int): B200Dataset =
    wrap(Native.matrixRankOneUpdate(h, leftRowNum, leftColNum, right.h, rightRowNum, rightColNum, blkSize))
  def transpose(): B200Dataset = wrap(Native.transpose(h))
  def t(): B200Dataset = transpose()
  def addScalar(alpha: Double): B200Dataset = wrap(Native.addScalar(h, alpha))
  def multiplyScalar(alpha: Double): B200Dataset = wrap(Native.multiplyScalar(h, alpha))
  def power(alpha: Double): B200Dataset = wrap(Native.power(h, alpha))
  override def finalize(): Unit = Native.matrixFree(h)
}
